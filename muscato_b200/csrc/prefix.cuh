// prefix.cuh -- device-wide exclusive prefix sum (uint32 in, uint32/uint64 out).
// This is the one grouping primitive of the hot path: counting-sort style segmented
// grouping (count -> exclusive scan -> scatter) replaces the reference's GNU sorts
// (cmd/muscato/main.go:237-304, 318-385, 453-463) wherever a group id is already known.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 4;  // one 16-byte load per thread and round
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

// One launch that fills up to kMaxFillJobs device buffers with a byte value each (replaces a train of
// cudaMemsetAsync calls: every memset is its own engine hand-over on the stream).
constexpr int kMaxFillJobs = 20;
struct FillJob {
  void* ptr[kMaxFillJobs];
  unsigned long long bytes[kMaxFillJobs];  // multiples of 16 (buffers are over-allocated accordingly)
  unsigned int value[kMaxFillJobs];        // 32-bit pattern
  int n;
};

__global__ void __launch_bounds__(256) fill_buffers_kernel(const FillJob job) {
  pdl_enter();
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

// ---------------------------------------------------------------------------------------------
// Single-launch exclusive scan on a grid that is resident as a whole (one block per SM).  Block b
// owns a contiguous slice: it sums the slice, publishes the sum, waits at a grid-wide barrier
// (a counter that only ever grows: launch k waits for k * gridDim.x arrivals, so it never has to
// be reset), derives its offset from the sums of the blocks before it and scans the slice.  The
// input is read twice with coalesced 16-byte loads; for the array sizes of this pipeline
// (10^6..10^8 counters) that beats a decoupled look-back chain, whose per-tile hand-over latency
// dominates below ~10^7 elements.
// ---------------------------------------------------------------------------------------------
constexpr int kBigItems = 16;  // elements per thread and round on the large-array path

template <typename OutT>
__global__ void __launch_bounds__(kScanThreads) scan_resident_kernel(const uint32_t* __restrict__ in,
                                                                     const unsigned long long* n_ptr, uint64_t n_host,
                                                                     OutT* __restrict__ out, int write_end,
                                                                     unsigned long long* __restrict__ total_out,
                                                                     uint64_t* block_sums, unsigned long long* barrier,
                                                                     unsigned long long barrier_target) {
  pdl_enter();
  __shared__ uint64_t warp_sums[kScanThreads / 32];
  __shared__ uint64_t s_total, s_off;
  const uint64_t n = scan_count(n_ptr, n_host);
  const uint64_t G = gridDim.x, b = blockIdx.x;
  const uint64_t per = (((n + G - 1) / G) + 3) & ~3ull;  // slice length, multiple of 4
  const uint64_t lo = min(n, b * per), hi = min(n, lo + per);

  // Small slices (up to 16 elements per thread, i.e. n <= ~1.2 M on 148 SMs): every thread keeps
  // its 16 contiguous elements in registers across the grid barrier -- the input is read once.
  constexpr int kRegItems = 16;
  const bool in_regs = per <= (uint64_t)kScanThreads * kRegItems;
  uint32_t rv[kRegItems];
  uint64_t ex_in_block = 0;

  // phase 1: slice sum
  uint64_t s = 0;
  if (in_regs) {
    const uint64_t i0 = lo + (uint64_t)threadIdx.x * kRegItems;
#pragma unroll
    for (int q = 0; q < kRegItems; q += 4) {
      const uint64_t i = i0 + q;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (i + 4 <= hi) {
        v = *reinterpret_cast<const uint4*>(in + i);
      } else if (i < hi) {
        v.x = in[i];
        if (i + 1 < hi) v.y = in[i + 1];
        if (i + 2 < hi) v.z = in[i + 2];
      }
      rv[q] = v.x; rv[q + 1] = v.y; rv[q + 2] = v.z; rv[q + 3] = v.w;
      s += (uint64_t)v.x + v.y + v.z + v.w;
    }
    ex_in_block = block_exclusive_scan_u64(s, &s_total, warp_sums);
  } else {
    // kBigItems contiguous elements per thread and round: four 16-byte loads in flight per thread (one block per
    // SM has to keep HBM busy on its own)
    for (uint64_t i0 = lo + (uint64_t)threadIdx.x * kBigItems; i0 < hi; i0 += (uint64_t)kScanThreads * kBigItems) {
      if (i0 + kBigItems <= hi) {
        uint4 v[kBigItems / 4];
#pragma unroll
        for (int q = 0; q < kBigItems / 4; q++) v[q] = *reinterpret_cast<const uint4*>(in + i0 + 4 * q);
#pragma unroll
        for (int q = 0; q < kBigItems / 4; q++) s += (uint64_t)v[q].x + v[q].y + v[q].z + v[q].w;
      } else {
        for (uint64_t j = i0; j < hi; j++) s += in[j];
      }
    }
    (void)block_exclusive_scan_u64(s, &s_total, warp_sums);
  }
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile uint64_t*>(block_sums + b) = s_total;
    __threadfence();
    atomicAdd(barrier, 1ull);
    while (*reinterpret_cast<volatile unsigned long long*>(barrier) < barrier_target) {}
    __threadfence();
  }
  __syncthreads();

  // phase 2: offset of this slice = sum of the earlier slices (gridDim.x <= kScanThreads)
  {
    const uint64_t v = threadIdx.x < G ? __ldcg(block_sums + threadIdx.x) : 0ull;
    const uint64_t ex = block_exclusive_scan_u64(v, &s_total, warp_sums);
    if (threadIdx.x == b) s_off = ex;
    __syncthreads();
    if (b == 0 && threadIdx.x == 0) {
      *total_out = s_total;
      if (write_end) out[n] = (OutT)s_total;
    }
  }
  uint64_t run0 = s_off;
  __syncthreads();

  // phase 3: scan of the slice
  if (in_regs) {
    const uint64_t i0 = lo + (uint64_t)threadIdx.x * kRegItems;
    uint64_t run = run0 + ex_in_block;
#pragma unroll
    for (int q = 0; q < kRegItems; q++) {
      if (i0 + q < hi) out[i0 + q] = (OutT)run;
      run += rv[q];
    }
    return;
  }
  for (uint64_t base = lo; base < hi; base += (uint64_t)kScanThreads * kBigItems) {
    const uint64_t i0 = base + (uint64_t)threadIdx.x * kBigItems;
    uint32_t v[kBigItems];
    const bool whole = i0 + kBigItems <= hi;
    if (whole) {
#pragma unroll
      for (int q = 0; q < kBigItems / 4; q++) {
        const uint4 w = *reinterpret_cast<const uint4*>(in + i0 + 4 * q);
        v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < kBigItems; q++) v[q] = i0 + q < hi ? in[i0 + q] : 0u;
    }
    uint64_t t = 0;
#pragma unroll
    for (int q = 0; q < kBigItems; q++) t += v[q];
    uint64_t run = run0 + block_exclusive_scan_u64(t, &s_total, warp_sums);
    if (whole) {
      // slices start at multiples of 4 elements and rounds at multiples of kBigItems: 16-byte aligned vector stores
      if (sizeof(OutT) == 8) {
#pragma unroll
        for (int q = 0; q < kBigItems; q += 2) {
          const uint64_t a0 = run, a1 = run + v[q];
          run = a1 + v[q + 1];
          *reinterpret_cast<ulonglong2*>(out + i0 + q) = make_ulonglong2(a0, a1);
        }
      } else {
#pragma unroll
        for (int q = 0; q < kBigItems; q += 4) {
          const uint64_t a0 = run, a1 = a0 + v[q], a2 = a1 + v[q + 1], a3 = a2 + v[q + 2];
          run = a3 + v[q + 3];
          *reinterpret_cast<uint4*>(out + i0 + q) = make_uint4((uint32_t)a0, (uint32_t)a1, (uint32_t)a2, (uint32_t)a3);
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < kBigItems; q++) {
        if (i0 + q < hi) out[i0 + q] = (OutT)run;
        run += v[q];
      }
    }
    run0 += s_total;
    __syncthreads();  // s_total / warp_sums are reused by the next round
  }
}

}  // namespace msc
