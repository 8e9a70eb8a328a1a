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
  const uint64_t n = min((uint64_t)*n_ptr, cap);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = m[i];
    if (v.w <= __ldg(best + v.x) + mmtol) atomicAdd(rcount + v.x, 1u);
  }
}

__global__ void __launch_bounds__(256) combine_scatter_kernel(const uint4* __restrict__ m,
                                                              const unsigned long long* __restrict__ n_ptr, uint64_t cap,
                                                              const uint32_t* __restrict__ best, uint32_t mmtol,
                                                              const uint32_t* __restrict__ rstart,
                                                              uint32_t* __restrict__ rfill, uint4* __restrict__ out) {
  const uint64_t n = min((uint64_t)*n_ptr, cap);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = m[i];
    if (v.w <= __ldg(best + v.x) + mmtol) out[__ldg(rstart + v.x) + atomicAdd(rfill + v.x, 1u)] = v;
  }
}

// Number of key groups whose passing-pair count exceeds MaxMatches (the only groups for
// which qinsert / "first" truncation, cmd/muscato_confirm/main.go:233-242 and :424-448, can
// drop anything).
__global__ void __launch_bounds__(256) overflow_count_kernel(const uint32_t* __restrict__ pass_cnt, uint64_t n_slots,
                                                             unsigned long long max_matches,
                                                             const unsigned long long* __restrict__ n_pass,
                                                             unsigned long long* __restrict__ n_over) {
  if (*n_pass <= max_matches) return;  // no group can exceed MaxMatches
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t over = (i < n_slots && (unsigned long long)pass_cnt[i] > max_matches) ? 1u : 0u;
  over = __reduce_add_sync(0xffffffffu, over);
  if ((threadIdx.x & 31u) == 0 && over) atomicAdd(n_over, (unsigned long long)over);
}

}  // namespace msc
