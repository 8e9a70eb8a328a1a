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

// One launch that fills up to six device buffers with a byte value each (replaces a train of
// cudaMemsetAsync calls: every memset is its own engine hand-over on the stream).
struct FillJob {
  void* ptr[6];
  unsigned long long bytes[6];  // multiples of 16 (buffers are over-allocated accordingly)
  unsigned int value[6];        // 32-bit pattern
  int n;
};

__global__ void __launch_bounds__(256) fill_buffers_kernel(const FillJob job) {
  for (int j = 0; j < job.n; j++) {
    uint4* p = reinterpret_cast<uint4*>(job.ptr[j]);
    const unsigned long long n16 = job.bytes[j] >> 4;
    const unsigned int v = job.value[j];
    const uint4 val = make_uint4(v, v, v, v);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (unsigned long long)gridDim.x * blockDim.x)
      p[i] = val;
  }
}

// Element count of a scan: a host value, or (n_ptr != nullptr) a device counter clamped to n.
__device__ __forceinline__ uint64_t scan_count(const unsigned long long* n_ptr, uint64_t n) {
  if (!n_ptr) return n;
  const unsigned long long v = *n_ptr;
  return v < n ? v : n;
}

// Pass 1: per-tile sums (grid-stride over tiles, so the launch does not depend on n).
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t* __restrict__ in,
                                                               const unsigned long long* n_ptr, uint64_t n_host,
                                                               uint64_t* __restrict__ tile_sums) {
  __shared__ uint64_t warp_sums[kScanThreads / 32];
  __shared__ uint64_t total;
  const uint64_t n = scan_count(n_ptr, n_host);
  const uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint64_t base = tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++)
      if (base + i < n) s += in[base + i];
    block_exclusive_scan_u64(s, &total, warp_sums);
    if (threadIdx.x == 0) tile_sums[tile] = total;
    __syncthreads();
  }
}

// Pass 2: one block scans the tile sums in place (exclusive); writes the grand total.
__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets(uint64_t* __restrict__ tile_sums,
                                                                  const unsigned long long* n_ptr, uint64_t n_host,
                                                                  uint64_t* __restrict__ grand_total) {
  __shared__ uint64_t warp_sums[kScanThreads / 32];
  __shared__ uint64_t total;
  const uint64_t ntiles = (scan_count(n_ptr, n_host) + kScanTile - 1) / kScanTile;
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
__global__ void __launch_bounds__(kScanThreads) scan_apply(const uint32_t* __restrict__ in,
                                                           const unsigned long long* n_ptr, uint64_t n_host,
                                                           const uint64_t* __restrict__ tile_offs,
                                                           OutT* __restrict__ out, int write_end) {
  __shared__ uint64_t warp_sums[kScanThreads / 32];
  __shared__ uint64_t total;
  const uint64_t n = scan_count(n_ptr, n_host);
  const uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
  if (n == 0 && write_end && blockIdx.x == 0 && threadIdx.x == 0) out[0] = (OutT)0;
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint64_t base = tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
      v[i] = base + i < n ? in[base + i] : 0u;
      s += v[i];
    }
    uint64_t run = tile_offs[tile] + block_exclusive_scan_u64(s, &total, warp_sums);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
      if (base + i < n) out[base + i] = (OutT)run;
      run += v[i];
      if (write_end && base + i + 1 == n) out[n] = (OutT)run;
    }
    __syncthreads();
  }
}

}  // namespace msc
