// prefix.cuh -- device-wide exclusive prefix sum (uint32 in, uint32/uint64 out).
// This is the one grouping primitive of the hot path: counting-sort style segmented
// grouping (count -> exclusive scan -> scatter) replaces the reference's GNU sorts
// (cmd/muscato/main.go:237-304, 318-385, 453-463) wherever a group id is already known.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
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
// Single-launch exclusive scan (decoupled look-back).  Tiles are handed out by an atomic ticket,
// so a tile only ever waits on tiles held by blocks that are already running.  Each tile
// publishes one 64-bit descriptor {status:2 | epoch:14 | value:48}: first its aggregate, then
// its inclusive prefix; warp 0 of the block walks back over the predecessors 32 at a time.  The
// epoch makes descriptors of earlier scans invisible, so the descriptor array is never cleared
// between scans; the ticket / completion counters reset themselves when the last block leaves.
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kDescValueMask = (1ull << 48) - 1ull;
constexpr uint32_t kScanEpochs = 1u << 14;
constexpr uint64_t kDescAggregate = 1ull, kDescInclusive = 2ull;

__device__ __forceinline__ uint64_t desc_pack(uint64_t status, uint32_t epoch, uint64_t value) {
  return (status << 62) | ((uint64_t)epoch << 48) | (value & kDescValueMask);
}

template <typename OutT>
__global__ void __launch_bounds__(kScanThreads) scan_onepass_kernel(const uint32_t* __restrict__ in,
                                                                    const unsigned long long* n_ptr, uint64_t n_host,
                                                                    OutT* __restrict__ out, int write_end,
                                                                    unsigned long long* __restrict__ total_out,
                                                                    uint64_t* desc, unsigned long long* state,
                                                                    uint32_t epoch) {
  __shared__ uint64_t warp_sums[kScanThreads / 32];
  __shared__ uint64_t s_total, s_tile, s_excl;
  const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const uint64_t n = scan_count(n_ptr, n_host);
  const uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
  if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    if (write_end) out[0] = (OutT)0;
    *total_out = 0ull;
  }
  while (true) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&state[0], 1ull);
    __syncthreads();
    const uint64_t tile = s_tile;
    if (tile >= ntiles) break;
    const uint64_t base = tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
      v[i] = base + i < n ? in[base + i] : 0u;
      s += v[i];
    }
    const uint64_t in_block = block_exclusive_scan_u64(s, &s_total, warp_sums);
    if (wid == 0) {
      const uint64_t agg = s_total;
      uint64_t excl = 0;
      if (tile > 0) {
        if (lane == 0) *reinterpret_cast<volatile uint64_t*>(desc + tile) = desc_pack(kDescAggregate, epoch, agg);
        int64_t back = (int64_t)tile - 1;
        while (true) {
          const int64_t p = back - (int64_t)lane;
          uint64_t status = kDescInclusive, val = 0;
          if (p >= 0) {
            uint64_t d;
            do {
              d = *reinterpret_cast<volatile uint64_t*>(desc + p);
            } while (((d >> 48) & (kScanEpochs - 1)) != epoch || (d >> 62) == 0);
            status = d >> 62;
            val = d & kDescValueMask;
          }
          const unsigned incl = __ballot_sync(0xffffffffu, status == kDescInclusive);
          const int first = incl ? __ffs(incl) - 1 : 31;
          uint64_t c = (int)lane <= first ? val : 0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
          excl += c;
          if (incl) break;
          back -= 32;
        }
      }
      if (lane == 0) {
        *reinterpret_cast<volatile uint64_t*>(desc + tile) = desc_pack(kDescInclusive, epoch, excl + agg);
        s_excl = excl;
        if (tile + 1 == ntiles) {
          *total_out = excl + agg;
          if (write_end) out[n] = (OutT)(excl + agg);
        }
      }
    }
    __syncthreads();
    uint64_t run = s_excl + in_block;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
      if (base + i < n) out[base + i] = (OutT)run;
      run += v[i];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&state[1], 1ull) == (unsigned long long)gridDim.x - 1ull) {
      state[0] = 0ull;  // every block has stopped taking tickets: ready for the next scan
      state[1] = 0ull;
    }
  }
}

}  // namespace msc
