// prefix.cuh -- device-wide exclusive prefix sum (uint32 in, uint32/uint64 out).
// This is the one grouping primitive of the hot path: counting-sort style segmented
// grouping (count -> exclusive scan -> scatter) replaces the reference's GNU sorts
// (cmd/muscato/main.go:237-304, 318-385, 453-463) wherever a group id is already known.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t* total, uint64_t* warp_sums) {
  const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((int)lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint64_t ws = lane < (kScanThreads / 32) ? warp_sums[lane] : 0;
    uint64_t winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if ((int)lane >= o) winc += t;
    }
    if (lane < (kScanThreads / 32)) warp_sums[lane] = winc - ws;
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  return warp_sums[wid] + inc - v;
}

// Pass 1: per-tile sums.
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t* __restrict__ in, uint64_t n,
                                                               uint64_t* __restrict__ tile_sums) {
  __shared__ uint64_t warp_sums[kScanThreads / 32];
  __shared__ uint64_t total;
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++)
    if (base + i < n) s += in[base + i];
  block_exclusive_scan_u64(s, &total, warp_sums);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Pass 2: one block scans the tile sums in place (exclusive); writes the grand total.
__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets(uint64_t* __restrict__ tile_sums, uint64_t ntiles,
                                                                  uint64_t* __restrict__ grand_total) {
  __shared__ uint64_t warp_sums[kScanThreads / 32];
  __shared__ uint64_t total;
  uint64_t carry = 0;
  for (uint64_t start = 0; start < ntiles; start += kScanThreads) {
    const uint64_t i = start + threadIdx.x;
    const uint64_t v = i < ntiles ? tile_sums[i] : 0;
    const uint64_t ex = block_exclusive_scan_u64(v, &total, warp_sums);
    if (i < ntiles) tile_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

// Pass 3: rescan each tile with its offset.  out[n] (one past the end) receives the total
// when write_end is set, so out can serve directly as a CSR offsets array.
template <typename OutT>
__global__ void __launch_bounds__(kScanThreads) scan_apply(const uint32_t* __restrict__ in, uint64_t n,
                                                           const uint64_t* __restrict__ tile_offs,
                                                           OutT* __restrict__ out, int write_end) {
  __shared__ uint64_t warp_sums[kScanThreads / 32];
  __shared__ uint64_t total;
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    v[i] = base + i < n ? in[base + i] : 0u;
    s += v[i];
  }
  uint64_t run = tile_offs[blockIdx.x] + block_exclusive_scan_u64(s, &total, warp_sums);
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) out[base + i] = (OutT)run;
    run += v[i];
    if (write_end && base + i + 1 == n) out[n] = (OutT)run;
  }
}

}  // namespace msc
