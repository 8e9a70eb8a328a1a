// common.cuh -- shared device helpers for the muscato_b200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace msc {

// ---------------------------------------------------------------------------
// 2-bit sequence layout.
//   base i of a stream lives in 64-bit word i>>5 at bits [2*(i&31), 2*(i&31)+2).
//   codes: A=0 C=1 T=2 G=3 ((ascii>>1)&3); any other byte is "X": code 0 plus a bit in
//   the X plane, which uses the same spacing (bit 2*(i&31) of word i>>5) so that the
//   same funnel shifts apply to both planes.
// ---------------------------------------------------------------------------
constexpr uint64_t kEvenBits = 0x5555555555555555ull;

// Bits 0, 2, 4, ... 30 of x packed into the low 16 bits.
__host__ __device__ __forceinline__ uint32_t compress_even32(uint32_t x) {
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0f0f0f0fu;
  x = (x | (x >> 4)) & 0x00ff00ffu;
  x = (x | (x >> 8)) & 0x0000ffffu;
  return x;
}

__host__ __device__ __forceinline__ uint64_t low_bases_mask(int n) {  // n in [0,32]
  return n >= 32 ? ~0ull : ((1ull << (2 * n)) - 1ull);
}

// 32 bases starting at base index `base` (reads word wi and wi+1: buffers are padded).
__device__ __forceinline__ uint64_t extract32(const uint64_t* __restrict__ w, uint64_t base) {
  const uint64_t wi = base >> 5;
  const unsigned sh = (unsigned)(base & 31u) * 2u;
  const uint64_t lo = __ldg(w + wi);
  if (sh == 0) return lo;
  const uint64_t hi = __ldg(w + wi + 1);
  return (lo >> sh) | (hi << (64u - sh));
}

// ---------------------------------------------------------------------------
// Window-key fingerprint.  A W<=32 window is 2W bits (key) plus its X mask (xm, same
// spacing).  An X-free window's fingerprint is the key itself (+1, so that it is never 0): exact
// and free; the table's home bucket comes from one multiplicative hash (table_home_bucket).  Windows
// containing X mix the mask in.  fp only has to be free of false negatives: the confirm kernel
// re-checks the window bases exactly (cmd/muscato_confirm/main.go:382-393 requires byte
// equality), so the rare collisions (an X window with an X-free one, the all-G 32-mer with the
// all-A one) only cost a rejected pair.  fp==0 is reserved for "empty slot".
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t z) {
  z ^= z >> 33;
  z *= 0xff51afd7ed558ccdull;
  z ^= z >> 33;
  z *= 0xc4ceb9fe1a85ec53ull;
  z ^= z >> 33;
  return z;
}

__host__ __device__ __forceinline__ uint64_t key_fp(uint64_t key, uint64_t xm) {
  uint64_t z = xm == 0 ? key + 1ull : fmix64((key ^ (xm * 0xD6E8FEB86659FD93ull)) + 0x9E3779B97F4A7C15ull);
  return z ? z : 1ull;
}

// Wide windows (32 < W <= 64): the window is two words (bases 0..31 and 32..W-1, each with its X
// mask); the fingerprint is a hash of all four, so it is never exact and the confirm kernel
// always re-checks the window bases.
__host__ __device__ __forceinline__ uint64_t key_fp_wide(uint64_t k0, uint64_t k1, uint64_t xm0, uint64_t xm1) {
  uint64_t z = fmix64(k0 + 0x9E3779B97F4A7C15ull) ^ (fmix64(k1 ^ 0xD6E8FEB86659FD93ull) * 0xC2B2AE3D27D4EB4Full);
  if (xm0 | xm1) z ^= fmix64(xm0 * 0xD6E8FEB86659FD93ull + xm1 * 0x9E3779B97F4A7C15ull + 1ull);
  z = fmix64(z);
  return z ? z : 1ull;
}
// What a wide key contributes to the 32-bit Bloom hash besides its first 16 bases.
__host__ __device__ __forceinline__ uint32_t wide_khi(uint64_t k0, uint64_t k1) {
  return (uint32_t)(k0 >> 32) ^ ((uint32_t)k1 * 0x9E3779B1u) ^ ((uint32_t)(k1 >> 32) * 0xC2B2AE35u);
}

// ---------------------------------------------------------------------------
// Blocked Bloom front: one 64-bit word per key (a single 8-byte load per probed target
// position), 2 bits in each 32-bit half.
//
// Addressing is LOCALITY AWARE: the 32-byte sector (4 words) of a key is chosen by the
// minimiser of the key's first P = min(W,16) bases -- the smallest of its wn = P-m+1 m-mers,
// lexicographic with the later base more significant, on a relabelled alphabet (every 2-bit code
// XOR 1, so that poly-A is not the smallest m-mer) -- so the W-mers of consecutive target
// positions, which share their minimiser for (wn+1)/2 positions on average, probe the same sector
// and a warp that probes 32 consecutive positions touches ~32*2/(wn+1) sectors instead of 32.
// The word inside the sector and the bit positions come from a cheap 32-bit hash of the whole
// key.  Keys whose window contains X (xm != 0) are addressed by their fingerprint instead (no
// locality; rare).  The filter only has to be free of false negatives: build (build_keys_insert_kernel)
// and scan use the same functions below.
// ---------------------------------------------------------------------------
struct BloomGeom {
  int lg_words;    // log2(number of 64-bit words), 10..32
  int lg_blk;      // log2(words per minimiser-addressed block): 2 = one 32-byte sector (filter resident in the L2:
                   // the L1/L2 sector count is what a probe costs), 4 = one 128-byte line (filter beyond the L2: HBM
                   // delivers whole lines, measured 118..128 B per random access, so the line is the unit to share)
  int m;           // minimiser length in bases, 1..16
  int wn;          // m-mers per key that compete: P - m + 1, 1..8
  uint32_t xr;     // alphabet relabelling of the key's low 32 bits: 0x55555555 cut to W bases
  // EXACT front (round 2): for W <= 15 and a key set so large that the Bloom words would take more room than one bit
  // per POSSIBLE key (4^W bits: 128 MB at W = 15), the front is that bitmap -- bit x = the key itself, no hash, no
  // false positives (a window with X sets/tests a hashed bit: false positives only).  The scan walks the database
  // 2^lg_pass times; pass q tests only the positions whose bit lies in slice q of the bitmap (x >> (lg_words + 6 -
  // lg_pass) == q), so the bits a pass touches (<= 32 MB) stay in the L2 while the table lines stream past them.
  int direct;      // 1 = exact bitmap over the key space
  int lg_pass;     // log2(passes of the scan), direct only
};

// Fingerprint-addressed variant (keys with X).
__host__ __device__ __forceinline__ uint32_t bloom_mask_lo(uint64_t h) {
  return (1u << (unsigned)(h & 31u)) | (1u << (unsigned)((h >> 5) & 31u));
}
__host__ __device__ __forceinline__ uint32_t bloom_mask_hi(uint64_t h) {
  return (1u << (unsigned)((h >> 10) & 31u)) | (1u << (unsigned)((h >> 15) & 31u));
}

// 32-bit hash of an X-free key: prex = low 32 bits of the key ^ xr, khi = its high 32 bits
// (K32: the key has at most 16 bases, khi == 0).
template <bool K32>
__host__ __device__ __forceinline__ uint32_t bloom_hash32(uint32_t prex, uint32_t khi) {
  uint32_t x = prex;
  if (!K32) x ^= khi * 0x85EBCA6Bu;
  x *= 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  return x;
}

// The two half masks from the hash: two bits in the low half (a 10-bit pattern index, which the
// scan serves from a shared-memory table), the same pair rotated in the high half.
__host__ __device__ __forceinline__ uint32_t bloom_pattern(uint32_t i) {  // i in [0, 1024)
  return (1u << (i & 31u)) | (1u << ((i >> 5) & 31u));
}
__host__ __device__ __forceinline__ void bloom_masks32(uint32_t h, uint32_t& mlo, uint32_t& mhi) {
  mlo = bloom_pattern((h >> 2) & 1023u);
  const unsigned r = (h >> 12) & 31u;
  mhi = (mlo << r) | (mlo >> ((32u - r) & 31u));
}

// Sector of an X-free key.  m-mer j of prex is brought to the top of a 32-bit word by a left
// shift -- written as a multiplication by mul[j] = 1 << (32 - 2m - 2j) so that it issues on the
// FMA pipe (IMAD) instead of the ALU pipe the rest of the probe arithmetic saturates; the bits
// below the m-mer (earlier bases) only break ties between equal m-mers.
template <int WN>
__device__ __forceinline__ uint32_t bloom_min_mmer(uint32_t prex, const uint32_t* __restrict__ mul) {
  uint32_t v = prex * mul[0];
#pragma unroll
  for (int j = 1; j < WN; j++) v = min(v, prex * mul[j]);
  return v;
}
__host__ __device__ __forceinline__ uint32_t bloom_sector_of(uint32_t vmin, int m, int lg_words, int lg_blk) {
  return ((vmin >> (32u - 2u * (unsigned)m)) * 0x9E3779B1u) >> (32 + lg_blk - lg_words);  // lg_blocks = lg_words - lg_blk
}

__host__ __device__ __forceinline__ uint32_t bloom_sector_rt(uint32_t prex, int wn, int m, int lg_words, int lg_blk) {
  const unsigned s0 = 32u - 2u * (unsigned)m;
  uint32_t v = prex << s0;
  for (int j = 1; j < wn; j++) {
    const uint32_t c = prex << (s0 - 2u * j);
    v = c < v ? c : v;
  }
  return bloom_sector_of(v, m, lg_words, lg_blk);
}

// Word index and the two 32-bit half masks of a key (build side and the scan's X path; the
// scan's main path inlines the same arithmetic with WN as a template parameter).  key1 = bases
// 32.. of a wide window (0 for W <= 32); xm = OR of the X masks of all its words.
__host__ __device__ __forceinline__ void bloom_locate(uint64_t key, uint64_t xm, uint64_t fp, int W,
                                                      const BloomGeom& g, uint64_t& widx, uint32_t& mlo,
                                                      uint32_t& mhi, uint64_t key1 = 0ull) {
  if (g.direct) {
    // bit index in [0, 2^(lg_words + 6)): the key (4^W <= 2^(lg_words + 6)), or a hash of the fingerprint for an X window
    const uint64_t x = xm == 0 ? key : fp >> (64 - (g.lg_words + 6));
    widx = x >> 6;
    mlo = (x & 32u) ? 0u : 1u << (unsigned)(x & 31u);
    mhi = (x & 32u) ? 1u << (unsigned)(x & 31u) : 0u;
    return;
  }
  if (xm == 0) {
    const uint32_t prex = (uint32_t)key ^ g.xr;
    const uint32_t h = W <= 16 ? bloom_hash32<true>(prex, 0u)
                               : bloom_hash32<false>(prex, W <= 32 ? (uint32_t)(key >> 32) : wide_khi(key, key1));
    const uint32_t sec = bloom_sector_rt(prex, g.wn, g.m, g.lg_words, g.lg_blk);
    widx = ((uint64_t)sec << g.lg_blk) | (uint64_t)(h >> (32 - g.lg_blk));
    bloom_masks32(h, mlo, mhi);
  } else {
    widx = fp >> (64 - g.lg_words);
    mlo = bloom_mask_lo(fp);
    mhi = bloom_mask_hi(fp);
  }
}
// ---------------------------------------------------------------------------
// Key table: open addressing over BUCKETS that are exactly one 128-byte L2 line.
//
// Measured on B200 (profiles/README.md, round 2): a 32-byte sector miss brings the whole 128-byte line
// from HBM (dram__bytes per random 4..32-byte access = 118..128 B, cudaLimitMaxL2FetchGranularity
// ignored), so every random access costs one line whatever it reads.  A bucket therefore holds
// everything a look-up needs in ONE line:
//     bytes   0.. 39  fp[5]    64-bit fingerprints, 0 = empty
//     bytes  48..127  rec[5]   uint4 {item0, rmx0, start, cnt}: first (read, window) item of the key
//                              group, that read's record word (length | nmiss << 11 | hasX << 31) and
//                              the CSR range [start, start + cnt) of the group's FURTHER members
// A key lives in the first free slot of the first bucket of its probe sequence that is not full (home
// bucket from one multiplicative hash, then linear); there are no deletions, so a bucket with a free
// slot ends every look-up.  slot id = 5 * bucket + position.  The bucket count is partitions x 2^k
// (not a power of two): the home bucket is umulhi(hash, n_buckets), its partition home >> k.
// ---------------------------------------------------------------------------
constexpr int kBucketSlots = 5;
constexpr int kBucketBytes = 128;
constexpr int kBucketRecOff = 48;

struct TableGeom {
  uint64_t n_buckets;   // partitions << lg_bpp
  int lg_bpp;           // log2(buckets per partition)
  uint32_t n_parts;
};

__host__ __device__ __forceinline__ uint64_t table_home_bucket(uint64_t fp, uint64_t n_buckets) {
#ifdef __CUDA_ARCH__
  return __umul64hi(fp * 0x9E3779B97F4A7C15ull, n_buckets);
#else
  return (uint64_t)(((unsigned __int128)(fp * 0x9E3779B97F4A7C15ull) * n_buckets) >> 64);
#endif
}

// One 256-bit read-only global load (sm_100: LDG.E.256) of a 32-byte aligned record.
__device__ __forceinline__ void ldg256(const void* p, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
// The same through L2 only (.cg): for records that other SMs are updating with atomics.
__device__ __forceinline__ void ldcg256(const void* p, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ uint64_t ldcg64(const void* p) {
  uint64_t v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// L2 eviction priorities (createpolicy; SASS: the policy travels in the memory descriptor).  evict_last: the slice of
// the exact front a scan pass keeps probing; evict_first: table lines and candidate records that are touched once.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint32_t ldg32_hint(const uint32_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
// 256-bit read-only load of a line that will not be needed again (LDG.E.NA.EFL2.256)
__device__ __forceinline__ void ldg256_stream(const void* p, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
// 64-bit OR without a return value whose line keeps the given L2 priority (REDG.E.OR.64 with a policy descriptor).
__device__ __forceinline__ void red_or64_hint(unsigned long long* p, unsigned long long v, uint64_t pol) {
  asm volatile("red.global.or.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint4 ldg128_last_use(const uint4* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg128_hint(void* p, uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg64_hint(void* p, uint2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg32_hint(void* p, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// Request a line into the L2 without a destination register.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The same through the TMA engine: `bytes` (multiple of 16) from a 16-byte aligned address, no register, no LSU slot.
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ const uint8_t* bucket_ptr(const uint8_t* tab, uint64_t b) { return tab + b * (uint64_t)kBucketBytes; }
__device__ __forceinline__ uint8_t* bucket_ptr(uint8_t* tab, uint64_t b) { return tab + b * (uint64_t)kBucketBytes; }
__device__ __forceinline__ const uint4* slot_rec_ptr(const uint8_t* tab, uint64_t slot) {
  const uint64_t b = slot / kBucketSlots;
  return reinterpret_cast<const uint4*>(tab + b * (uint64_t)kBucketBytes + kBucketRecOff) + (slot - b * kBucketSlots);
}
__device__ __forceinline__ uint4* slot_rec_ptr(uint8_t* tab, uint64_t slot) {
  const uint64_t b = slot / kBucketSlots;
  return reinterpret_cast<uint4*>(tab + b * (uint64_t)kBucketBytes + kBucketRecOff) + (slot - b * kBucketSlots);
}

// The five fingerprints of a bucket (read-only path: the table is final).
__device__ __forceinline__ void bucket_load_fps(const uint8_t* bp, uint64_t q[kBucketSlots]) {
  ldg256(bp, q[0], q[1], q[2], q[3]);
  q[4] = __ldg(reinterpret_cast<const unsigned long long*>(bp + 32));
}

// Position of fp among a bucket's fingerprints: 0..4 = found, 5 = not here but the bucket has a free
// slot (the key is not in the table), 6 = bucket full, go on.
__device__ __forceinline__ int bucket_probe(uint64_t fp, const uint64_t q[kBucketSlots]) {
  int r = 6;
#pragma unroll
  for (int s = kBucketSlots - 1; s >= 0; s--) {
    if (q[s] == 0ull) r = 5;
  }
#pragma unroll
  for (int s = kBucketSlots - 1; s >= 0; s--) {
    if (q[s] == fp) r = s;
  }
  return r;
}

// Look-up.  Returns slot (= 5 * bucket + position) or -1.
__device__ __forceinline__ int64_t table_find(const uint8_t* __restrict__ tab, uint64_t n_buckets, uint64_t fp) {
  uint64_t b = table_home_bucket(fp, n_buckets);
  while (true) {
    uint64_t q[kBucketSlots];
    bucket_load_fps(bucket_ptr(tab, b), q);
    const int r = bucket_probe(fp, q);
    if (r < kBucketSlots) return (int64_t)(b * kBucketSlots + (uint64_t)r);
    if (r == kBucketSlots) return -1;
    b = b + 1 == n_buckets ? 0 : b + 1;
  }
}

// Programmatic dependent launch (PDL): every kernel of the pipeline starts with pdl_enter().  The
// launch_dependents half lets the NEXT kernel of the stream be launched (its CTAs scheduled as
// resources free up) while this grid is still running; the wait half blocks until the PREVIOUS
// grid has completed and its memory is visible -- so data dependencies are exactly those of a plain
// in-order stream, only the launch latency between two kernels is overlapped.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP / SYNCS).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Warp-aggregated append: every calling lane (any divergent subset) reserves one slot.
__device__ __forceinline__ unsigned long long warp_agg_inc(unsigned long long* counter) {
  const unsigned active = __activemask();
  const int leader = __ffs(active) - 1;
  const unsigned lane = threadIdx.x & 31u;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(active));
  base = __shfl_sync(active, base, leader);
  return base + (unsigned long long)__popc(active & ((1u << lane) - 1u));
}

// Warp-aggregated counter bump without a return value (compiles to a fire-and-forget RED).
__device__ __forceinline__ void warp_agg_count(unsigned long long* counter) {
  const unsigned active = __activemask();
  if ((int)(threadIdx.x & 31u) == __ffs(active) - 1) atomicAdd(counter, (unsigned long long)__popc(active));
}

// First index i in [lo, hi) with a[i] > v  (a ascending).
template <typename T>
__device__ __forceinline__ uint64_t upper_bound_dev(const T* __restrict__ a, uint64_t lo, uint64_t hi, T v) {
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (__ldg(a + mid) <= v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

}  // namespace msc
