// pack.cuh -- kernel family (1a): ASCII -> 2-bit packing of reads and targets on the device.
// Replaces the per-byte text handling of the reference (reads: cmd/muscato_screen/main.go:165-190,
// cmd/muscato_window_reads/main.go:100-126; targets: cmd/muscato_screen/main.go:440-452).
#pragma once
#include "common.cuh"

namespace msc {

// code / X flag of one ASCII base.  A=0x41 C=0x43 G=0x47 T=0x54 -> (c>>1)&3 = 0,1,3,2.
__device__ __forceinline__ void base_code(uint32_t c, uint32_t& code, uint32_t& isx) {
  const bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
  isx = ok ? 0u : 1u;
  code = ok ? ((c >> 1) & 3u) : 0u;
}

// 16 ASCII bases held in a uint4 -> 32 packed bits (+ 32 X-plane bits).
__device__ __forceinline__ void pack16(const uint4 v, uint32_t& bits, uint32_t& xbits) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  bits = 0;
  xbits = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
#pragma unroll
    for (int b = 0; b < 4; b++) {
      uint32_t code, isx;
      base_code((w[q] >> (8 * b)) & 0xffu, code, isx);
      const int sh = 2 * (4 * q + b);
      bits |= code << sh;
      xbits |= isx << sh;
    }
  }
}

// Reads: one thread per (read, word).  Row r of `words` / `xplane` has `stride` words;
// len_flags[r] = length | (has X ? 1<<31 : 0) (zero-initialised by the caller).
__global__ void __launch_bounds__(256) pack_reads_kernel(const uint8_t* __restrict__ ascii,
                                                         const uint64_t* __restrict__ offs, uint64_t n_reads,
                                                         int stride, uint64_t* __restrict__ words,
                                                         uint64_t* __restrict__ xplane,
                                                         uint32_t* __restrict__ len_flags) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_reads * (uint64_t)stride) return;
  const uint64_t r = idx / (uint64_t)stride;
  const int w = (int)(idx - r * (uint64_t)stride);
  const uint64_t o0 = offs[r];
  const int L = (int)(offs[r + 1] - o0);
  const int b0 = w * 32;
  const int n = min(32, L - b0);
  uint64_t bits = 0, xb = 0;
  const uint8_t* src = ascii + o0 + b0;
  for (int i = 0; i < n; i++) {
    uint32_t code, isx;
    base_code(__ldg(src + i), code, isx);
    bits |= (uint64_t)code << (2 * i);
    xb |= (uint64_t)isx << (2 * i);
  }
  words[idx] = bits;
  xplane[idx] = xb;
  uint32_t lf = (w == 0) ? (uint32_t)L : 0u;
  if (xb) lf |= 0x80000000u;
  if (lf) atomicOr(len_flags + r, lf);
}

// Targets: the ASCII stream is the concatenation of all targets (no separators); one
// thread packs 32 consecutive bases with two aligned 16-byte loads.  xplane must be
// zero-initialised; only words containing X are written.  xsum bit w = word w has X.
__global__ void __launch_bounds__(256) pack_targets_kernel(const uint8_t* __restrict__ ascii, uint64_t n_bases,
                                                           uint64_t* __restrict__ words, uint64_t n_words_alloc,
                                                           uint64_t* __restrict__ xplane,
                                                           uint32_t* __restrict__ xsum) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words_alloc) return;
  const uint64_t b0 = w * 32;
  uint64_t bits = 0, xb = 0;
  if (b0 + 32 <= n_bases) {
    const uint4* src = reinterpret_cast<const uint4*>(ascii + b0);
    uint32_t lo, xlo, hi, xhi;
    pack16(__ldg(src), lo, xlo);
    pack16(__ldg(src + 1), hi, xhi);
    bits = (uint64_t)lo | ((uint64_t)hi << 32);
    xb = (uint64_t)xlo | ((uint64_t)xhi << 32);
  } else if (b0 < n_bases) {
    const int n = (int)(n_bases - b0);
    for (int i = 0; i < n; i++) {
      uint32_t code, isx;
      base_code(__ldg(ascii + b0 + i), code, isx);
      bits |= (uint64_t)code << (2 * i);
      xb |= (uint64_t)isx << (2 * i);
    }
  }
  words[w] = bits;  // words past the stream end are zero padding
  if (xb) {
    xplane[w] = xb;
    atomicOr(xsum + (w >> 5), 1u << (unsigned)(w & 31u));
  }
}

}  // namespace msc
