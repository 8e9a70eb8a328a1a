// scan.cuh -- kernel (2): streaming scan of the 2-bit target database.
//
// Replaces search/processSeq/checkWin/harvest (cmd/muscato_screen/main.go:220-480): the
// reference rolls NumHash buzhash32 hashes over every target base and probes one
// BloomSize-bit array per window; here every target position's W-mer IS its key (W<=32),
// obtained by a funnel shift of two adjacent packed words, and ONE probe of a blocked Bloom
// word (8 bytes) per position covers all windows.  Positions that pass go to the exact key
// table; hits are compacted as (table slot, global target position) with warp-aggregated
// atomics.  Window/target-boundary and position-0 rules (processSeq :294-365) are applied
// by the confirm kernel, which knows the gene of each candidate.
//
// Data movement: target tiles (256 words = 8192 bases + halo) are staged into shared memory
// by the TMA engine (1-D cp.async.bulk + mbarrier, double buffered); a persistent grid walks
// the tiles.  The Bloom/table probes are scattered 8-byte reads served from L2 when the
// filter fits (it is sized 32-64 bits per key) and from HBM sectors otherwise.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int kScanBlock = 256;
constexpr int kTileWords = 256;                       // one 32-base word per thread per tile
constexpr int kTileCopyWords = kTileWords + 2;        // halo word + 1 (byte count multiple of 16)
constexpr int kTileSmemWords = kTileWords + 8;
constexpr int kStageCap = 128;                        // per-warp candidate staging (entries)

struct ScanArgs {
  const uint64_t* tg_words;
  const uint64_t* tg_x;
  const uint32_t* xsum;
  uint64_t n_bases;
  uint64_t n_tiles;
  const uint2* bloom;
  int lg_bloom;
  const uint64_t* tab_fp;
  int lg_slots;
  uint2* cand;
  unsigned long long cand_cap;
  unsigned long long* n_cand;
  unsigned long long* n_bloom_pass;
  int W;
};

// Extract the W-mer that starts at base j (0..31) of the word pair (lo, hi).
__device__ __forceinline__ uint64_t window_at(uint64_t lo, uint64_t hi, unsigned j, uint64_t kmask) {
  return (j == 0 ? lo : ((lo >> (2 * j)) | (hi << (64 - 2 * j)))) & kmask;
}

// Two phases per tile, so that the scattered probes never serialise behind divergent hit handling:
//   phase 1  every thread tests its 32 positions against the Bloom front (4 batches of 8
//            independent 8-byte loads) and keeps a 32-bit pass mask;
//   phase 2  the warp compacts its passes into a shared-memory queue and drains it 32 at a time:
//            lane i re-derives the key of queued position i with two shuffles, looks it up in the
//            exact table, and the warp appends the found (slot, position) pairs with ONE atomic.
__global__ void __launch_bounds__(kScanBlock) scan_targets_kernel(const ScanArgs a) {
  __shared__ alignas(128) uint64_t tile[2][kTileSmemWords];
  __shared__ alignas(8) uint64_t bar[2];
  __shared__ uint16_t queue[kScanBlock / 32][1024];
  __shared__ uint2 stage[kScanBlock / 32][kStageCap];  // found (slot, position) pairs, flushed when nearly full
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;
  constexpr uint32_t kBytes = kTileCopyWords * sizeof(uint64_t);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const uint64_t kmask = low_bases_mask(a.W);
  uint64_t t = blockIdx.x;
  if (tid == 0 && t < a.n_tiles) {
    mbar_arrive_expect_tx(&bar[0], kBytes);
    bulk_copy_g2s(tile[0], a.tg_words + t * kTileWords, kBytes, &bar[0]);
  }
  uint32_t phases = 0;
  int buf = 0;
  uint32_t n_pass = 0;
  uint16_t* q = queue[warp];
  uint2* st = stage[warp];
  uint32_t n_st = 0;  // staged candidates of this warp (warp-uniform)
  // One global atomic per flush instead of one per drain round: same-address atomics serialise.
  auto flush_stage = [&]() {
    unsigned long long out0 = 0;
    if (lane == 0) out0 = atomicAdd(a.n_cand, (unsigned long long)n_st);
    out0 = __shfl_sync(0xffffffffu, out0, 0);
    for (uint32_t i = lane; i < n_st; i += 32)
      if (out0 + i < a.cand_cap) a.cand[out0 + i] = st[i];
    __syncwarp();
    n_st = 0;
  };
  for (; t < a.n_tiles; t += gridDim.x) {
    const uint64_t tn = t + gridDim.x;
    if (tid == 0 && tn < a.n_tiles) {
      mbar_arrive_expect_tx(&bar[buf ^ 1], kBytes);
      bulk_copy_g2s(tile[buf ^ 1], a.tg_words + tn * kTileWords, kBytes, &bar[buf ^ 1]);
    }
    mbar_wait(&bar[buf], (phases >> buf) & 1u);
    phases ^= 1u << buf;

    const uint64_t lo = tile[buf][tid];
    const uint64_t hi = tile[buf][tid + 1];
    const uint64_t w = t * kTileWords + (uint64_t)tid;
    const uint64_t gbase = w * 32ull;
    uint32_t mask = 0;
    uint64_t xlo = 0, xhi = 0;
    if (gbase < a.n_bases) {
      const int npos = (int)min((uint64_t)32, a.n_bases - gbase);
      // X summary bits of word w and w+1 (xsum is padded).
      const uint32_t xs0 = __ldg(a.xsum + (w >> 5));
      const uint32_t xs1 = __ldg(a.xsum + ((w + 1) >> 5));
      const bool anyx = ((xs0 >> (unsigned)(w & 31u)) | (xs1 >> (unsigned)((w + 1) & 31u))) & 1u;
      if (!anyx) {
#pragma unroll
        for (int jb = 0; jb < 32; jb += 8) {
          uint64_t fp[8];
          uint2 bw[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            fp[i] = key_fp(window_at(lo, hi, jb + i, kmask), 0ull);
            bw[i] = __ldg(a.bloom + bloom_index(fp[i], a.lg_bloom));
          }
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const uint32_t mlo = bloom_mask_lo(fp[i]), mhi = bloom_mask_hi(fp[i]);
            if (((bw[i].x & mlo) == mlo) & ((bw[i].y & mhi) == mhi)) mask |= 1u << (jb + i);
          }
        }
      } else {
        // The word (or its successor) contains X: fold the X mask into the key.
        xlo = __ldg(a.tg_x + w);
        xhi = __ldg(a.tg_x + w + 1);
        for (int j = 0; j < 32; j++) {
          const uint64_t fp = key_fp(window_at(lo, hi, j, kmask), window_at(xlo, xhi, j, kmask));
          const uint2 bw = __ldg(a.bloom + bloom_index(fp, a.lg_bloom));
          const uint32_t mlo = bloom_mask_lo(fp), mhi = bloom_mask_hi(fp);
          if (((bw.x & mlo) == mlo) & ((bw.y & mhi) == mhi)) mask |= 1u << j;
        }
      }
      if (npos < 32) mask &= (1u << npos) - 1u;
    }

    // ---- phase 2: warp-cooperative drain of the Bloom passes -------------------------------
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total) {
      n_pass += cnt;
      uint32_t at = incl - cnt;
      uint32_t m = mask;
      while (m) {
        const uint32_t j = __ffs(m) - 1;
        m &= m - 1;
        q[at++] = (uint16_t)((lane << 5) | j);
      }
      __syncwarp();
      const uint64_t wbase = (t * kTileWords + (uint64_t)(warp * 32)) * 32ull;  // first position of this warp's words
      for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t idx = base + lane;
        const bool active = idx < total;
        const uint32_t e = active ? q[idx] : 0u;
        const unsigned src = e >> 5, j = e & 31u;
        const uint64_t l = __shfl_sync(0xffffffffu, lo, src), h = __shfl_sync(0xffffffffu, hi, src);
        const uint64_t xl = __shfl_sync(0xffffffffu, xlo, src), xh = __shfl_sync(0xffffffffu, xhi, src);
        int64_t slot = -1;
        if (active) slot = table_find(a.tab_fp, a.lg_slots, key_fp(window_at(l, h, j, kmask), window_at(xl, xh, j, kmask)));
        const unsigned found = __ballot_sync(0xffffffffu, slot >= 0);
        if (found) {
          if (slot >= 0) st[n_st + __popc(found & ((1u << lane) - 1u))] = make_uint2((uint32_t)slot, (uint32_t)(wbase + e));
          n_st += __popc(found);
          __syncwarp();
          if (n_st > kStageCap - 32) flush_stage();
        }
      }
      __syncwarp();
    }
    __syncthreads();  // all reads of tile[buf] are done before it is refilled
    buf ^= 1;
  }
  if (n_st) flush_stage();
  n_pass = __reduce_add_sync(0xffffffffu, n_pass);
  if (lane == 0 && n_pass) atomicAdd(a.n_bloom_pass, (unsigned long long)n_pass);
}

}  // namespace msc
