// readprep.cuh -- device-side prepReads: sort the raw reads bytewise and
// collapse duplicates, feeding the key-table build directly.
//
// Replaces, for the sequence column, `muscato_prep_reads | sort | muscato_uniqify`
// (cmd/muscato/main.go:152-221): non-ACGT -> X and truncation to MaxReadLength
// (cmd/muscato_prep_reads/main.go:33-44, :67-69), skipping reads shorter than MinReadLength
// (:59-62), the LC_ALL=C sort of the `seq\tname` lines (the sequence is the major key: '\t' sorts
// below every base letter, so a proper prefix sorts first) and the run-length collapse of equal
// sequences (cmd/muscato_uniqify/main.go:113-135).  Names are joined by the host from the
// permutation this returns (they need a per-group bytewise sort of their own).
//
// Method: every read becomes a fixed-width key of MRL 4-bit symbols (0 = past the end, A=1 C=2
// G=3 T=4 X=5, i.e. bytewise order with prefix-first), stored as byte planes (plane b holds
// symbols 2b, 2b+1 of every read); a stable LSD radix sort with 8-bit digits permutes read
// indices plane by plane (histogram -> single-launch scan -> stable scatter using
// __match_any_sync ranks); group heads are the positions whose key differs from the previous one.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int kRadixThreads = 256;
constexpr int kRadixItems = 8;  // items per thread and chunk
constexpr int kRadixChunk = kRadixThreads * kRadixItems;

__device__ __forceinline__ uint32_t prep_symbol(uint32_t c) {  // bytewise rank of the subx'ed base
  return c == 'A' ? 1u : c == 'C' ? 2u : c == 'G' ? 3u : c == 'T' ? 4u : 5u;
}

// One thread per (raw read, plane byte).  planes[b * n + i] = (sym(2b) << 4) | sym(2b+1) of read i
// truncated to max_len; reads shorter than min_len get 0xFF planes (they sort to the end and are
// cut off) and keep[i] = 0.
__global__ void __launch_bounds__(256) prep_encode_kernel(const uint8_t* __restrict__ ascii,
                                                          const uint64_t* __restrict__ offs, uint64_t n, int max_len,
                                                          int min_len, int n_planes, uint8_t* __restrict__ planes,
                                                          uint64_t* __restrict__ key64, uint32_t* __restrict__ keep,
                                                          unsigned long long* __restrict__ n_kept) {
  pdl_enter();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t k = 0;
  if (i < n) {
    const uint64_t o = offs[i];
    const int Lraw = (int)(offs[i + 1] - o);
    const bool kept = Lraw >= min_len;  // cmd/muscato_prep_reads/main.go:59-62 (tested before truncation)
    const int L = min(Lraw, max_len);
    uint64_t k64 = 0;  // planes 0..7 (the first 16 symbols), most significant first: the tie test's prefix
    for (int b = 0; b < n_planes; b++) {
      uint32_t v = 0xFFu;
      if (kept) {
        const int s0 = 2 * b, s1 = 2 * b + 1;
        const uint32_t hi = s0 < L ? prep_symbol(__ldg(ascii + o + s0)) : 0u;
        const uint32_t lo = s1 < L ? prep_symbol(__ldg(ascii + o + s1)) : 0u;
        v = (hi << 4) | lo;
      }
      planes[(uint64_t)b * n + i] = (uint8_t)v;
      if (b < 8) k64 |= (uint64_t)v << (56 - 8 * b);
    }
    // (planes beyond n_planes stay 0 in k64: equal for all reads)
    key64[i] = k64;
    keep[i] = kept ? 1u : 0u;
    k = kept ? 1u : 0u;
  }
  k = __reduce_add_sync(0xffffffffu, k);
  __shared__ uint32_t s_k[8];
  if ((threadIdx.x & 31u) == 0) s_k[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; w++) t += s_k[w];
    if (t) atomicAdd(n_kept, (unsigned long long)t);
  }
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ idx, uint64_t n) {
  pdl_enter();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (uint32_t)i;
}

// Radix pass, step 1: per-chunk digit histogram, written digit-major: hist[d * n_chunks + chunk].
__global__ void __launch_bounds__(kRadixThreads) radix_hist_kernel(const uint32_t* __restrict__ idx,
                                                                   const uint8_t* __restrict__ plane, uint64_t n,
                                                                   uint32_t n_chunks, uint32_t* __restrict__ hist) {
  pdl_enter();
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kRadixChunk;
#pragma unroll
  for (int t = 0; t < kRadixItems; t++) {
    const uint64_t i = base + (uint64_t)t * kRadixThreads + threadIdx.x;
    if (i < n) atomicAdd(&sh[__ldg(plane + __ldg(idx + i))], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * n_chunks + blockIdx.x] = sh[threadIdx.x];
}

// Radix pass, step 2 (after the exclusive scan of hist): stable scatter.  Items of a chunk are
// taken 256 at a time in their current order; the rank of an item among the equal digits of its
// sub-tile is (equal digits in earlier warps) + (equal digits in lower lanes of its warp).
__global__ void __launch_bounds__(kRadixThreads) radix_scatter_kernel(const uint32_t* __restrict__ idx_in,
                                                                      const uint8_t* __restrict__ plane, uint64_t n,
                                                                      uint32_t n_chunks,
                                                                      const uint32_t* __restrict__ offsets,
                                                                      uint32_t* __restrict__ idx_out) {
  pdl_enter();
  __shared__ uint32_t base[256];     // running output position of every digit for this chunk
  __shared__ uint32_t wh[8][256];    // per-warp digit counts of the current sub-tile -> exclusive positions
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  base[threadIdx.x] = offsets[(uint64_t)threadIdx.x * n_chunks + blockIdx.x];
  const uint64_t cbase = (uint64_t)blockIdx.x * kRadixChunk;
  for (int t = 0; t < kRadixItems; t++) {
#pragma unroll
    for (int q = 0; q < 8; q++) wh[q][threadIdx.x] = 0;
    __syncthreads();
    const uint64_t i = cbase + (uint64_t)t * kRadixThreads + threadIdx.x;
    const bool live = i < n;
    uint32_t id = 0, d = 0;
    if (live) {
      id = __ldg(idx_in + i);
      d = __ldg(plane + id);
    }
    // lanes that are past the end take a digit of their own (256 + lane) so that they never pair up
    const unsigned peers = __match_any_sync(0xffffffffu, live ? d : 256u + lane);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (live && rank == 0) wh[w][d] = __popc(peers);
    __syncthreads();
    {
      const uint32_t dg = threadIdx.x;  // one digit per thread: exclusive prefix over the warps
      uint32_t acc = base[dg];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const uint32_t c = wh[q][dg];
        wh[q][dg] = acc;
        acc += c;
      }
      base[dg] = acc;
    }
    __syncthreads();
    if (live) idx_out[wh[w][d] + rank] = id;
    __syncthreads();
  }
}

// Tie fix after a sort on the first `n_pre` planes only: runs of equal prefix (almost always
// duplicates or a handful of reads that share their first 2 * n_pre bases) are put in full-key
// order by the thread that owns the run's first position (insertion sort, keys compared plane by
// plane from n_pre on).  Runs longer than kMaxTieRun are left alone and reported through
// `too_long`; the caller then falls back to a sort over all planes.
constexpr int kMaxTieRun = 64;

__device__ __forceinline__ bool prep_prefix_equal(const uint64_t* __restrict__ key64, uint32_t a, uint32_t b) {
  return key64[a] == key64[b];  // the first 8 planes (n_pre == 8), packed by prep_encode_kernel
}
__device__ __forceinline__ bool prep_suffix_less(const uint8_t* __restrict__ planes, uint64_t n, int n_pre, int n_planes,
                                                 uint32_t a, uint32_t b) {
  for (int p = n_pre; p < n_planes; p++) {
    const uint8_t ka = planes[(uint64_t)p * n + a], kb = planes[(uint64_t)p * n + b];
    if (ka != kb) return ka < kb;
  }
  return false;
}

__global__ void __launch_bounds__(256) prep_tiefix_kernel(uint32_t* __restrict__ idx, const uint8_t* __restrict__ planes,
                                                          const uint64_t* __restrict__ key64, uint64_t n,
                                                          const unsigned long long* __restrict__ n_kept_ptr,
                                                          int n_pre, int n_planes,
                                                          unsigned long long* __restrict__ too_long) {
  pdl_enter();
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_kept = *n_kept_ptr;
  if (j >= n_kept) return;
  if (j > 0 && prep_prefix_equal(key64, idx[j], idx[j - 1])) return;  // not the first of its run
  uint64_t len = 1;
  while (j + len < n_kept && len <= (uint64_t)kMaxTieRun && prep_prefix_equal(key64, idx[j + len], idx[j])) len++;
  if (len == 1) return;
  if (len > (uint64_t)kMaxTieRun) {
    atomicMax(too_long, (unsigned long long)len);
    return;
  }
  for (uint64_t a = 1; a < len; a++) {  // stable insertion sort on the key suffix
    const uint32_t v = idx[j + a];
    uint64_t b = a;
    while (b > 0 && prep_suffix_less(planes, n, n_pre, n_planes, v, idx[j + b - 1])) {
      idx[j + b] = idx[j + b - 1];
      b--;
    }
    idx[j + b] = v;
  }
}

// head[j] = 1 when sorted position j starts a new sequence (compared over all key planes).
__global__ void __launch_bounds__(256) prep_heads_kernel(const uint32_t* __restrict__ idx,
                                                         const uint8_t* __restrict__ planes,
                                                         const uint64_t* __restrict__ key64, uint64_t n,
                                                         const unsigned long long* __restrict__ n_kept_ptr,
                                                         int n_planes, uint32_t* __restrict__ head) {
  pdl_enter();
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_kept = *n_kept_ptr;
  if (j >= n_kept) return;
  uint32_t h = 1;
  if (j > 0) {
    const uint32_t a = idx[j], b = idx[j - 1];
    h = key64[a] != key64[b] ? 1u : 0u;  // planes 0..7 in one compare
    for (int p = 8; p < n_planes && !h; p++)
      if (planes[(uint64_t)p * n + a] != planes[(uint64_t)p * n + b]) h = 1;
  }
  head[j] = h;
}

// group id (exclusive scan of head, +head) -> group_start[], unique read lengths.
__global__ void __launch_bounds__(256) prep_groups_kernel(const uint32_t* __restrict__ idx,
                                                          const uint32_t* __restrict__ head,
                                                          const uint32_t* __restrict__ head_scan,
                                                          const unsigned long long* __restrict__ n_kept_ptr,
                                                          const uint64_t* __restrict__ raw_offs, int max_len,
                                                          const unsigned long long* __restrict__ n_unique_ptr,
                                                          uint32_t* __restrict__ group_start,
                                                          uint32_t* __restrict__ ulen) {
  pdl_enter();
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_kept = *n_kept_ptr;
  if (j == 0) group_start[*n_unique_ptr] = (uint32_t)n_kept;  // sentinel: end of the last group
  if (j >= n_kept) return;
  if (head[j]) {
    const uint32_t u = head_scan[j];  // heads before j = index of this group
    group_start[u] = (uint32_t)j;
    const uint32_t r = idx[j];
    const uint64_t L = raw_offs[r + 1] - raw_offs[r];
    ulen[u] = (uint32_t)min(L, (uint64_t)max_len);
  }
}

// Copy the representative of every unique read into the context's read buffer (subx'ed, truncated).
__global__ void __launch_bounds__(256) prep_gather_kernel(const uint8_t* __restrict__ raw_ascii,
                                                          const uint64_t* __restrict__ raw_offs,
                                                          const uint32_t* __restrict__ idx,
                                                          const uint32_t* __restrict__ group_start,
                                                          const uint64_t* __restrict__ uoffs,
                                                          const unsigned long long* __restrict__ n_unique_ptr,
                                                          uint8_t* __restrict__ out_ascii) {
  pdl_enter();
  // one warp per unique read
  const uint64_t u = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (u >= *n_unique_ptr) return;
  const uint32_t r = idx[group_start[u]];
  const uint64_t src = raw_offs[r], dst = uoffs[u];
  const uint32_t L = (uint32_t)(uoffs[u + 1] - dst);
  for (uint32_t i = lane; i < L; i += 32) {
    const uint8_t c = raw_ascii[src + i];
    out_ascii[dst + i] = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : (uint8_t)'X';
  }
}

}  // namespace msc
