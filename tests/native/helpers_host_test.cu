// Host-side checks of the __host__ __device__ arithmetic the kernels share (muscato_b200/csrc/common.cuh): compiled by
// nvcc, run on the CPU by tests/test_native_helpers_cpu.py (no CUDA runtime call is made).
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../muscato_b200/csrc/common.cuh"

using namespace msc;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

#define CHECK(c)                                                         \
  do {                                                                   \
    if (!(c)) {                                                          \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);         \
      return 1;                                                          \
    }                                                                    \
  } while (0)

int main() {
  // compress_even32: bits 0, 2, ..., 30 -> the low 16 bits
  for (int it = 0; it < 100000; it++) {
    const uint32_t x = (uint32_t)rnd();
    uint32_t want = 0;
    for (int j = 0; j < 16; j++) want |= ((x >> (2 * j)) & 1u) << j;
    CHECK(compress_even32(x) == want);
  }
  // low_bases_mask
  CHECK(low_bases_mask(0) == 0ull && low_bases_mask(15) == (1ull << 30) - 1ull && low_bases_mask(32) == ~0ull);
  // key_fp: never 0, exact (key + 1) for X-free windows, deterministic with X
  for (int it = 0; it < 100000; it++) {
    const uint64_t key = rnd() & low_bases_mask(20), xm = (it & 7) ? 0ull : (rnd() & kEvenBits & low_bases_mask(20));
    const uint64_t fp = key_fp(key, xm);
    CHECK(fp != 0ull);
    if (xm == 0) CHECK(fp == key + 1ull);
    CHECK(fp == key_fp(key, xm));
  }
  // table_home_bucket: inside the table for bucket counts that are not powers of two
  for (int it = 0; it < 100000; it++) {
    const uint64_t nb = 1 + rnd() % 200000000ull;
    CHECK(table_home_bucket(rnd(), nb) < nb);
  }
  // exact front (BloomGeom::direct): an X-free key sets / tests exactly bit x = key of the map, and the slice a scan
  // launch derives from the word index (widx >> (lg_words - lg_pass)) is the key's top lg_pass bits (key >> pshift),
  // which is what the member masks of scan_direct.cuh select on; X windows land inside the map
  for (int W = 8; W <= 15; W++) {
    for (int lg_pass = 0; lg_pass <= 4; lg_pass++) {
      BloomGeom g{};
      g.direct = 1;
      g.lg_words = 2 * W - 6 > 10 ? 2 * W - 6 : 10;
      g.lg_pass = lg_pass;
      g.lg_blk = 2;
      g.m = 1;
      g.wn = 1;
      const unsigned pshift = (unsigned)(g.lg_words + 6 - g.lg_pass);
      for (int it = 0; it < 20000; it++) {
        const uint64_t key = rnd() & low_bases_mask(W);
        uint64_t widx;
        uint32_t mlo, mhi;
        bloom_locate(key, 0ull, key + 1ull, W, g, widx, mlo, mhi);
        const uint64_t mask64 = (uint64_t)mlo | ((uint64_t)mhi << 32);
        CHECK(widx == key >> 6 && mask64 == 1ull << (key & 63u));
        CHECK((widx >> (g.lg_words - g.lg_pass)) == (key >> pshift));
        CHECK(widx < (1ull << g.lg_words));
        // the 32-bit view the scan probes: word key >> 5, bit key & 31
        const uint32_t half = (key & 32u) ? mhi : mlo;
        CHECK(half == 1u << (key & 31u));
        const uint64_t xm = (rnd() & kEvenBits & low_bases_mask(W)) | 1ull;
        bloom_locate(key, xm, key_fp(key, xm), W, g, widx, mlo, mhi);
        CHECK(widx < (1ull << g.lg_words));
        CHECK(__builtin_popcountll((uint64_t)mlo | ((uint64_t)mhi << 32)) == 1);
      }
    }
  }
  // Bloom front: build side (bloom_locate) and the scan's inlined arithmetic agree on sector and masks
  for (int W : {15, 20, 32}) {
    BloomGeom g{};
    g.lg_words = 20;
    g.lg_blk = 2;
    const int P = W < 16 ? W : 16;
    g.m = 9;
    g.wn = P - g.m + 1;
    g.xr = 0x55555555u & (uint32_t)low_bases_mask(W);
    for (int it = 0; it < 20000; it++) {
      const uint64_t key = rnd() & low_bases_mask(W);
      uint64_t widx;
      uint32_t mlo, mhi;
      bloom_locate(key, 0ull, key + 1ull, W, g, widx, mlo, mhi);
      const uint32_t prex = (uint32_t)key ^ g.xr;
      const uint32_t h = W <= 16 ? bloom_hash32<true>(prex, 0u) : bloom_hash32<false>(prex, (uint32_t)(key >> 32));
      const uint32_t sec = bloom_sector_rt(prex, g.wn, g.m, g.lg_words, g.lg_blk);
      uint32_t a, b;
      bloom_masks32(h, a, b);
      CHECK(widx == (((uint64_t)sec << g.lg_blk) | (uint64_t)(h >> (32 - g.lg_blk))) && a == mlo && b == mhi);
      CHECK(widx < (1ull << g.lg_words) && __builtin_popcount(mlo) >= 1 && __builtin_popcount(mlo) <= 2);
    }
  }
  std::printf("ok\n");
  return 0;
}
