// Micro-benchmark: cost of the random-access primitives used by the key-table build on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics_bench atomics_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__host__ __device__ inline uint64_t mix(uint64_t z) { z ^= z >> 33; z *= 0xff51afd7ed558ccdull; z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ull; z ^= z >> 33; return z; }
__global__ void k_cas64(unsigned long long* t, uint64_t mask, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; if (i >= n) return;
  uint64_t fp = mix(i + 1) | 1; uint64_t s = (fp * 0x9E3779B97F4A7C15ull) >> 40 & mask;
  while (true) { unsigned long long c = atomicCAS(t + s, 0ull, fp); if (c == 0 || c == fp) break; s = (s + 1) & mask; }
}
__global__ void k_cas32(unsigned int* t, uint64_t mask, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; if (i >= n) return;
  uint64_t fp = mix(i + 1); uint32_t tag = (uint32_t)fp | 1u; uint64_t s = (fp * 0x9E3779B97F4A7C15ull) >> 40 & mask;
  while (true) { unsigned int c = atomicCAS(t + s, 0u, tag); if (c == 0 || c == tag) break; s = (s + 1) & mask; }
}
__global__ void k_or64(unsigned long long* t, uint64_t mask, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; if (i >= n) return;
  uint64_t fp = mix(i + 1); atomicOr(t + ((fp >> 20) & mask), 1ull << (fp & 63));
}
__global__ void k_store32(unsigned int* t, uint64_t mask, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; if (i >= n) return;
  uint64_t fp = mix(i + 1); t[(fp >> 20) & mask] = (unsigned)i;
}
__global__ void k_load64(const unsigned long long* t, uint64_t mask, uint64_t n, unsigned long long* out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; if (i >= n) return;
  uint64_t fp = mix(i + 1); unsigned long long v = t[(fp >> 20) & mask]; if (v == 12345) out[0] = v;
}
__global__ void k_add32(unsigned int* t, uint64_t mask, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; if (i >= n) return;
  uint64_t fp = mix(i + 1); atomicAdd(t + ((fp >> 20) & mask), 1u);
}
template <typename F> float timeit(F f, void* buf, size_t bytes, bool warm, void* flush, size_t fbytes) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float best = 1e9;
  for (int rep = 0; rep < 5; rep++) {
    if (!warm) { cudaMemset(buf, 0, bytes); cudaMemset(flush, 1, fbytes); } else { cudaMemset(flush, 1, fbytes); cudaMemset(buf, 0, bytes); }
    cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best * 1000.f;
}
int main() {
  const uint64_t n = 2000000; const int lg = 22; const uint64_t slots = 1ull << lg, mask = slots - 1;
  void *t, *flush; unsigned long long* out; cudaMalloc(&t, slots * 8); size_t fb = 512ull << 20; cudaMalloc(&flush, fb); cudaMalloc(&out, 8);
  unsigned g = (unsigned)((n + 255) / 256);
  for (int warm = 0; warm < 2; warm++) {
    printf("--- table %s in L2 (memset %s flush)\n", warm ? "warm" : "cold", warm ? "after" : "before");
    printf("cas64  2M keys into 4M slots (33 MB): %8.1f us\n", timeit([&] { k_cas64<<<g, 256>>>((unsigned long long*)t, mask, n); }, t, slots * 8, warm, flush, fb));
    printf("cas32  2M keys into 4M slots (16 MB): %8.1f us\n", timeit([&] { k_cas32<<<g, 256>>>((unsigned*)t, mask, n); }, t, slots * 4, warm, flush, fb));
    printf("or64   2M into 1M words (8 MB)      : %8.1f us\n", timeit([&] { k_or64<<<g, 256>>>((unsigned long long*)t, (1 << 20) - 1, n); }, t, 8 << 20, warm, flush, fb));
    printf("add32  2M into 4M words (16 MB)     : %8.1f us\n", timeit([&] { k_add32<<<g, 256>>>((unsigned*)t, mask, n); }, t, slots * 4, warm, flush, fb));
    printf("store32 2M into 4M words (16 MB)    : %8.1f us\n", timeit([&] { k_store32<<<g, 256>>>((unsigned*)t, mask, n); }, t, slots * 4, warm, flush, fb));
    printf("load64 2M from 4M words (33 MB)     : %8.1f us\n", timeit([&] { k_load64<<<g, 256>>>((unsigned long long*)t, mask, n, out); }, t, slots * 8, warm, flush, fb));
    printf("load64 20M from 1M words (8 MB)     : %8.1f us\n", timeit([&] { k_load64<<<(unsigned)((20000000 + 255) / 256), 256>>>((unsigned long long*)t, (1 << 20) - 1, 20000000, out); }, t, 8 << 20, warm, flush, fb));
  }
  printf("memset 33 MB: %8.1f us\n", timeit([&] { cudaMemsetAsync(t, 0, slots * 8); }, t, 8, true, flush, 8));
  return 0;
}
