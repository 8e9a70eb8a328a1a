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

// 4 ASCII bases in one 32-bit word -> 8 packed bits + 8 X-plane bits (SWAR, no per-byte loop).
// code = (byte>>1)&3; the byte is valid iff it equals LUT[code] with LUT = {A,C,T,G}, which
// __byte_perm evaluates for all four bytes at once.
__device__ __forceinline__ void pack4(uint32_t v, uint32_t& bits, uint32_t& xbits) {
  uint32_t c = (v >> 1) & 0x03030303u;
  uint32_t t = c | (c >> 4);                                  // nibble-pack the codes as a permute selector
  const uint32_t sel = (t & 0xffu) | ((t >> 8) & 0xff00u);
  const uint32_t expect = __byte_perm(0x47544341u, 0u, sel);  // bytes: 0:'A' 1:'C' 2:'T' 3:'G'
  const uint32_t diff = v ^ expect;
  xbits = 0u;
  if (diff) {  // some byte is not A/C/G/T (rare): clear its code, set its X bit
    const uint32_t nz = ((diff | ((diff & 0x7f7f7f7fu) + 0x7f7f7f7fu)) >> 7) & 0x01010101u;  // 1 per invalid byte
    c &= ~(nz * 3u);
    xbits = (nz * 0x01041040u) >> 24;
  }
  // gather the four 2-bit fields (bits 0, 8, 16, 24) into one byte: the partial products of the
  // multiplication land on disjoint bit pairs, the wanted ones in the top byte (IMAD: FMA pipe)
  bits = (c * 0x01041040u) >> 24;
}

// 16 ASCII bases held in a uint4 -> 32 packed bits (+ 32 X-plane bits).
__device__ __forceinline__ void pack16(const uint4 v, uint32_t& bits, uint32_t& xbits) {
  uint32_t b0, x0, b1, x1, b2, x2, b3, x3;
  pack4(v.x, b0, x0);
  pack4(v.y, b1, x1);
  pack4(v.z, b2, x2);
  pack4(v.w, b3, x3);
  bits = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
  xbits = x0 | (x1 << 8) | (x2 << 16) | (x3 << 24);
}

// Reads: one thread per (read, word); a block packs `reads_per_block` = 256/stride whole reads.
// Row r of `words` / `xplane` has `stride` words; len_flags[r] = length | (has X ? 1<<31 : 0).
// The block's reads are one contiguous byte range of the concatenated ASCII: it is staged into
// shared memory with coalesced 16-byte loads and each thread then picks its (arbitrarily
// aligned) 32 bytes out of shared memory with 32-bit loads + funnel shifts.
__global__ void __launch_bounds__(256) pack_reads_kernel(const uint8_t* __restrict__ ascii,
                                                         const uint64_t* __restrict__ offs, uint64_t n_reads,
                                                         int stride, int reads_per_block,
                                                         uint64_t* __restrict__ words, uint64_t* __restrict__ xplane,
                                                         uint32_t* __restrict__ len_flags) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t sm[];
  const uint64_t r0 = (uint64_t)blockIdx.x * (uint64_t)reads_per_block;
  const uint64_t r1 = min(r0 + (uint64_t)reads_per_block, n_reads);
  const uint64_t blk_lo = __ldg(offs + r0), blk_hi = __ldg(offs + r1);
  const uint64_t lo_al = blk_lo & ~15ull;
  const uint32_t nbytes = (uint32_t)(blk_hi - lo_al);
  for (uint32_t i = threadIdx.x * 16u; i < nbytes; i += 256u * 16u)
    *reinterpret_cast<uint4*>(sm + i) = __ldg(reinterpret_cast<const uint4*>(ascii + lo_al + i));
  __syncthreads();
  const uint32_t lr = threadIdx.x / (uint32_t)stride;
  const int w = (int)(threadIdx.x - lr * (uint32_t)stride);
  const uint64_t r = r0 + lr;
  if ((int)lr >= reads_per_block || r >= r1) return;
  const uint64_t o0 = __ldg(offs + r);
  const int L = (int)(__ldg(offs + r + 1) - o0);
  const int b0 = w * 32;
  const int n = min(32, L - b0);
  uint64_t bits = 0, xb = 0;
  if (n > 0) {
    const uint32_t a = (uint32_t)(o0 + (uint64_t)b0 - lo_al);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(sm + (a & ~3u));
    const unsigned sh = (a & 3u) * 8u;
    uint32_t v[9];
#pragma unroll
    for (int i = 0; i < 9; i++) v[i] = wp[i];
    uint4 q0, q1;
    q0.x = __funnelshift_r(v[0], v[1], sh);
    q0.y = __funnelshift_r(v[1], v[2], sh);
    q0.z = __funnelshift_r(v[2], v[3], sh);
    q0.w = __funnelshift_r(v[3], v[4], sh);
    q1.x = __funnelshift_r(v[4], v[5], sh);
    q1.y = __funnelshift_r(v[5], v[6], sh);
    q1.z = __funnelshift_r(v[6], v[7], sh);
    q1.w = __funnelshift_r(v[7], v[8], sh);
    uint32_t blo, xlo, bhi, xhi;
    pack16(q0, blo, xlo);
    pack16(q1, bhi, xhi);
    const uint64_t keep = low_bases_mask(n);
    bits = ((uint64_t)blo | ((uint64_t)bhi << 32)) & keep;
    xb = ((uint64_t)xlo | ((uint64_t)xhi << 32)) & keep;
  }
  const uint64_t idx = r * (uint64_t)stride + (uint64_t)w;
  words[idx] = bits;
  xplane[idx] = xb;
  uint32_t lf = (w == 0) ? (uint32_t)L : 0u;
  if (xb) lf |= 0x80000000u;
  if (lf) atomicOr(len_flags + r, lf);
}

// Offsets that arrive in device memory (msc_set_reads_device) are validated here instead of on
// the host: offs[0]==0, monotone, offs[n]==total, every length <= max_len.
__global__ void __launch_bounds__(256) validate_read_offsets_kernel(const uint64_t* __restrict__ offs, uint64_t n_reads,
                                                                    uint64_t total, uint64_t max_len,
                                                                    unsigned long long* __restrict__ bad) {
  pdl_enter();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_reads) return;
  const uint64_t a = offs[i], b = offs[i + 1];
  bool ok = b >= a && b - a <= max_len;
  if (i == 0) ok = ok && a == 0;
  if (i + 1 == n_reads) ok = ok && b == total;
  if (!ok) atomicOr(bad, 1ull);
}

// Targets: the ASCII stream is the concatenation of all targets (no separators); one
// thread packs 32 consecutive bases with two aligned 16-byte loads.  xplane must be
// zero-initialised; only words containing X are written.  xsum bit w = word w has X.
__global__ void __launch_bounds__(256) pack_targets_kernel(const uint8_t* __restrict__ ascii, uint64_t n_bases,
                                                           uint64_t* __restrict__ words, uint64_t n_words_alloc,
                                                           uint64_t* __restrict__ xplane,
                                                           uint32_t* __restrict__ xsum,
                                                           unsigned long long* __restrict__ any_x) {
  pdl_enter();
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
    atomicOr(any_x, 1ull);
  }
}

// Targets that arrive already packed (msc_set_targets_packed: the persistent 2-bit target cache):
// words / xplane were copied into place; this kernel zeroes the padding behind the stream, clears
// the bits of the last word past the stream's end and derives the per-word X summary bits.
__global__ void __launch_bounds__(256) packed_targets_finish_kernel(uint64_t* __restrict__ words, uint64_t* __restrict__ xplane,
                                                                    uint64_t n_bases, uint64_t n_words_alloc,
                                                                    uint32_t* __restrict__ xsum,
                                                                    unsigned long long* __restrict__ any_x) {
  pdl_enter();
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words_alloc) return;
  const uint64_t b0 = w * 32;
  if (b0 >= n_bases) {
    words[w] = 0ull;
    xplane[w] = 0ull;
    return;
  }
  uint64_t xb = xplane[w];
  if (b0 + 32 > n_bases) {
    const uint64_t keep = low_bases_mask((int)(n_bases - b0));
    words[w] &= keep;
    xb &= keep & kEvenBits;
    xplane[w] = xb;
  }
  if (xb) {
    atomicOr(xsum + (w >> 5), 1u << (unsigned)(w & 31u));
    atomicOr(any_x, 1ull);
  }
}

}  // namespace msc
