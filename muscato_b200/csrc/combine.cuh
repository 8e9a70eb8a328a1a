// combine.cuh -- the MMTol filter and per-read grouping of confirmed matches.
// Replaces writebest (cmd/muscato_combine_windows/main.go:36-60): per read keep the lines
// with nx <= min_nx + MMTol.  Grouping by read is a counting sort on the read id
// (count -> exclusive scan -> scatter), which also yields the reads_sorted order that
// `sort -u` (cmd/muscato/main.go:453-463) establishes on the first column.
#pragma once
#include "common.cuh"

namespace msc {

__global__ void __launch_bounds__(256) combine_count_kernel(const uint4* __restrict__ m,
                                                            const unsigned long long* __restrict__ n_ptr, uint64_t cap,
                                                            const uint32_t* __restrict__ best, uint32_t mmtol,
                                                            uint32_t* __restrict__ rcount) {
  pdl_enter();
  const uint64_t n = min((uint64_t)*n_ptr, cap);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = m[i];
    if (v.w <= __ldg(best + v.x) + mmtol) atomicAdd(rcount + v.x, 1u);
  }
}

__global__ void __launch_bounds__(256) combine_scatter_kernel(const uint4* __restrict__ m,
                                                              const unsigned long long* __restrict__ n_ptr, uint64_t cap,
                                                              const uint32_t* __restrict__ best, uint32_t mmtol,
                                                              uint32_t* __restrict__ rcursor, uint4* __restrict__ out) {
  pdl_enter();
  // rcursor[r] starts as a copy of the read's first output slot (rstart[r]) and is bumped per match: ONE random line
  // per kept match next to best[r], where a separate fill counter cost a second one
  const uint64_t n = min((uint64_t)*n_ptr, cap);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = m[i];
    if (v.w <= __ldg(best + v.x) + mmtol) out[atomicAdd(rcursor + v.x, 1u)] = v;
  }
}

// Order inside each read group: (gene, pos) ascending, so that the library's output is
// deterministic without a host-side sort.  Short groups (the norm) are sorted by one thread in
// registers; longer ones are queued for segment_rank_sort_kernel.
constexpr int kShortSegment = 8;

__device__ __forceinline__ bool match_less(const uint4& a, const uint4& b) {
  return a.y != b.y ? a.y < b.y : a.z < b.z;
}

__global__ void __launch_bounds__(256) segment_sort_short_kernel(uint4* __restrict__ m,
                                                                 const uint32_t* __restrict__ rstart, uint64_t n_reads,
                                                                 uint32_t* __restrict__ long_list,
                                                                 unsigned long long* __restrict__ n_long,
                                                                 uint32_t* __restrict__ mid_list,
                                                                 unsigned long long* __restrict__ n_mid) {
  pdl_enter();
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint32_t lo = rstart[r], hi = rstart[r + 1];
  const uint32_t n = hi - lo;
  if (n <= 1) return;
  if (n > kShortSegment) {  // queued: one warp per group up to 32 members, one block per larger group
    if (n <= 32u) mid_list[atomicAdd(n_mid, 1ull)] = (uint32_t)r;
    else long_list[atomicAdd(n_long, 1ull)] = (uint32_t)r;
    return;
  }
  uint4 v[kShortSegment];
#pragma unroll
  for (int i = 0; i < kShortSegment; i++)
    if (i < (int)n) v[i] = m[lo + i];
  // insertion sort with static indexing (odd-even transposition keeps v[] in registers)
#pragma unroll
  for (int pass = 0; pass < kShortSegment; pass++) {
#pragma unroll
    for (int i = pass & 1; i + 1 < kShortSegment; i += 2) {
      if (i + 1 < (int)n && match_less(v[i + 1], v[i])) {
        const uint4 t = v[i];
        v[i] = v[i + 1];
        v[i + 1] = t;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kShortSegment; i++)
    if (i < (int)n) m[lo + i] = v[i];
}

// One block per long read group.  Groups of up to kRankSmem members are sorted in shared memory
// (bitonic network); larger ones fall back to a rank sort in global memory: every element counts
// the elements that precede it (keys are unique inside a group after de-duplication) and is
// written at that rank into `scratch`, then the sorted segment is copied back.
constexpr int kRankThreads = 256;
constexpr int kRankSmem = 2048;

__global__ void __launch_bounds__(kRankThreads) segment_rank_sort_kernel(
    uint4* __restrict__ m, uint4* __restrict__ scratch, const uint32_t* __restrict__ rstart,
    const uint32_t* __restrict__ long_list, const unsigned long long* __restrict__ n_long,
    const uint32_t* __restrict__ mid_list, const unsigned long long* __restrict__ n_mid) {
  pdl_enter();
  __shared__ uint64_t keys[kRankSmem];
  __shared__ uint2 rest[kRankSmem];
  const uint64_t nl = *n_long, nm = *n_mid;
  // pass 1: groups of up to 32 members, one WARP per group: every lane holds one member and
  // counts the members that precede it with 32 shuffles (no shared memory, no block barrier)
  {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t gw = ((uint64_t)blockIdx.x * kRankThreads + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * kRankThreads) >> 5;
    for (uint64_t s = gw; s < nm; s += nw) {
      const uint32_t r = mid_list[s];
      const uint32_t lo = rstart[r], n = rstart[r + 1] - lo;
      uint4 a = make_uint4(0u, 0u, 0u, 0u);
      if (lane < n) a = m[lo + lane];
      const uint64_t k = lane < n ? (((uint64_t)a.y << 32) | (uint64_t)a.z) : ~0ull;
      uint32_t rank = 0;
      for (uint32_t j = 0; j < n; j++) rank += __shfl_sync(0xffffffffu, k, j) < k ? 1u : 0u;
      __syncwarp();
      if (lane < n) m[lo + rank] = a;
    }
  }
  // pass 2: larger groups, one block per group
  for (uint64_t s = blockIdx.x; s < nl; s += gridDim.x) {
    const uint32_t r = long_list[s];
    const uint32_t lo = rstart[r], hi = rstart[r + 1], n = hi - lo;
    if (n <= (uint32_t)kRankSmem) {
      // bitonic sort of (gene, pos) keys in shared memory (padded to a power of two with +inf keys)
      uint32_t np2 = 2;
      while (np2 < n) np2 <<= 1;
      for (uint32_t i = threadIdx.x; i < np2; i += kRankThreads) {
        if (i < n) {
          const uint4 a = m[lo + i];
          keys[i] = ((uint64_t)a.y << 32) | (uint64_t)a.z;
          rest[i] = make_uint2(a.x, a.w);
        } else {
          keys[i] = ~0ull;
          rest[i] = make_uint2(0u, 0u);
        }
      }
      __syncthreads();
      for (uint32_t k = 2; k <= np2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t t = threadIdx.x; t < np2 / 2; t += kRankThreads) {
            const uint32_t i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));  // j is a power of two
            const bool up = (i & k) == 0;
            const uint64_t x = keys[i], y = keys[i + j];
            if ((x > y) == up) {
              keys[i] = y;
              keys[i + j] = x;
              const uint2 r0 = rest[i];
              rest[i] = rest[i + j];
              rest[i + j] = r0;
            }
          }
          __syncthreads();
        }
      }
      for (uint32_t i = threadIdx.x; i < n; i += kRankThreads)
        m[lo + i] = make_uint4(rest[i].x, (uint32_t)(keys[i] >> 32), (uint32_t)keys[i], rest[i].y);
      __syncthreads();
    } else {
      for (uint32_t i = lo + threadIdx.x; i < hi; i += kRankThreads) {
        const uint4 a = m[i];
        uint32_t rank = 0;
        for (uint32_t j = lo; j < hi; j++) rank += match_less(m[j], a) ? 1u : 0u;
        scratch[lo + rank] = a;
      }
      __syncthreads();
      for (uint32_t i = lo + threadIdx.x; i < hi; i += kRankThreads) m[i] = scratch[i];
      __syncthreads();
    }
  }
}

// Non-match writer support (cmd/muscato_nonmatch/main.go:57-113 keeps the reads
// whose sequence is absent from column 1 of the results): a read matched iff its best mismatch
// count was ever set.  flag -> exclusive scan -> ordered scatter yields the unmatched read ids in
// reads_sorted order, which is the order of the non-match fastq.
__global__ void __launch_bounds__(256) nonmatch_flag_kernel(const uint32_t* __restrict__ best, uint64_t n_reads,
                                                            uint32_t no_match, uint32_t* __restrict__ flag) {
  pdl_enter();
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_reads) flag[r] = best[r] == no_match ? 1u : 0u;
}
__global__ void __launch_bounds__(256) nonmatch_scatter_kernel(const uint32_t* __restrict__ flag,
                                                               const uint32_t* __restrict__ pos, uint64_t n_reads,
                                                               uint32_t* __restrict__ list) {
  pdl_enter();
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_reads && flag[r]) list[pos[r]] = (uint32_t)r;
}

// Flag the key groups whose passing-pair count exceeds MaxMatches (mode-2 input).
__global__ void __launch_bounds__(256) overflow_flag_kernel(const uint32_t* __restrict__ pass_cnt, uint64_t n_slots,
                                                            unsigned long long max_matches,
                                                            uint8_t* __restrict__ slot_over) {
  pdl_enter();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_slots) slot_over[i] = (unsigned long long)pass_cnt[i] > max_matches ? 1 : 0;
}

// best[read] = min nx over a match list (used after the host merged truncated groups back in).
__global__ void __launch_bounds__(256) best_from_matches_kernel(const uint4* __restrict__ m, uint64_t n,
                                                                uint32_t* __restrict__ best) {
  pdl_enter();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMin(best + m[i].x, m[i].w);
}

// Number of counters above MaxMatches.  `cnt` is either the small hashed counter array of the confirm
// kernel (every counter is an upper bound for the key groups that map to it: none above the limit
// means no group can lose anything to qinsert / "first" truncation,
// cmd/muscato_confirm/main.go:233-242 and :424-448) or, in the re-run that follows, the exact
// per-slot counts.
__global__ void __launch_bounds__(256) overflow_count_kernel(const uint32_t* __restrict__ cnt, uint64_t n,
                                                             unsigned long long max_matches,
                                                             const unsigned long long* __restrict__ n_pass,
                                                             unsigned long long* __restrict__ n_over,
                                                             uint32_t* __restrict__ shard_flag) {
  pdl_enter();
  if (*n_pass <= max_matches) return;  // no group can exceed MaxMatches (nothing to read)
  uint32_t over = 0;
  const uint64_t n4 = n / 4;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(cnt) + i);
    over += ((unsigned long long)v.x > max_matches) + ((unsigned long long)v.y > max_matches) +
            ((unsigned long long)v.z > max_matches) + ((unsigned long long)v.w > max_matches);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3u)) over += (unsigned long long)cnt[n4 * 4 + threadIdx.x] > max_matches;
  over = __reduce_add_sync(0xffffffffu, over);
  if ((threadIdx.x & 31u) == 0 && over) {
    atomicAdd(n_over, (unsigned long long)over);
    // sharded targets (max_matches = MaxMatches / n_shards): the flag sits behind the per-read best
    // array and reaches every rank with the MIN all-reduce of that array
    if (shard_flag) *shard_flag = 0u;
  }
}

// Fingerprints of the key groups whose exact passing-pair count exceeds `thr` (sharded MaxMatches
// protocol: a fingerprint names a k-mer independently of the rank's table layout).
__global__ void __launch_bounds__(256) overflow_keys_kernel(const uint32_t* __restrict__ pass_cnt, const uint8_t* __restrict__ tab,
                                                            uint64_t n_slots, unsigned long long thr, uint64_t* __restrict__ out,
                                                            unsigned long long cap, unsigned long long* __restrict__ n_out) {
  pdl_enter();
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += (uint64_t)gridDim.x * blockDim.x) {
    if ((unsigned long long)pass_cnt[s] <= thr) continue;
    const uint64_t b = s / kBucketSlots;
    const uint64_t fp = reinterpret_cast<const uint64_t*>(tab + b * (uint64_t)kBucketBytes)[s - b * kBucketSlots];
    if (!fp) continue;
    const unsigned long long at = atomicAdd(n_out, 1ull);
    if (at < cap) out[at] = fp;
  }
}

// slot_over[slot] = 1 for every listed fingerprint that is a key of this table.
__global__ void __launch_bounds__(256) flag_keys_kernel(const uint64_t* __restrict__ fps, uint64_t n, const uint8_t* __restrict__ tab,
                                                        uint64_t n_buckets, uint8_t* __restrict__ slot_over) {
  pdl_enter();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t s = table_find(tab, n_buckets, fps[i]);
  if (s >= 0) slot_over[s] = 1;
}

// After the all-reduce: copy the flag into the counter block the host reads anyway.
__global__ void shard_flag_kernel(const uint32_t* __restrict__ flag, unsigned long long* __restrict__ out) {
  pdl_enter();
  if (threadIdx.x == 0) *out = *flag == 0u ? 1ull : 0ull;
}

}  // namespace msc
