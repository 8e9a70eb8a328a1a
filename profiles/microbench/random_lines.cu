// random_lines.cu -- how many RANDOM 128-byte lines per second does this GPU's HBM deliver, as a function of
// the footprint?  Every thread issues kInFlight independent 32-byte loads (one sector of a random line) per
// round; the addresses come from a hash of (thread, round).  Built and run by profiles/gpu_call*.sh:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o random_lines random_lines.cu && ./random_lines
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z ^= z >> 33; z *= 0xff51afd7ed558ccdull; z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ull; z ^= z >> 33;
  return z;
}

template <int IN_FLIGHT>
__global__ void __launch_bounds__(256) probe(const uint8_t* __restrict__ buf, uint64_t n_lines, int rounds, uint64_t seed,
                                             unsigned long long* sink) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t acc = 0;
  for (int r = 0; r < rounds; r++) {
    uint64_t v[IN_FLIGHT][4];
#pragma unroll
    for (int i = 0; i < IN_FLIGHT; i++) {
      const uint64_t h = mix(seed + tid * 0x9E3779B97F4A7C15ull + (uint64_t)(r * IN_FLIGHT + i));
      const uint64_t line = __umul64hi(h, n_lines);
      const uint8_t* p = buf + line * 128ull + ((h & 3ull) << 5);
      asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[i][0]), "=l"(v[i][1]), "=l"(v[i][2]), "=l"(v[i][3]) : "l"(p));
    }
#pragma unroll
    for (int i = 0; i < IN_FLIGHT; i++) acc ^= v[i][0] ^ v[i][1] ^ v[i][2] ^ v[i][3];
  }
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}

int main() {
  int dev = 0;
  cudaSetDevice(dev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  const uint64_t sizes_mb[] = {64, 128, 256, 512, 1024, 4096, 16384, 65536};
  uint8_t* buf = nullptr;
  const uint64_t max_bytes = 65536ull << 20;
  if (cudaMalloc(&buf, max_bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 1, max_bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("%s, %d SMs; per size: in-flight per thread, blocks/SM -> random 128-byte lines/s (x128 B)\n", prop.name, prop.multiProcessorCount);
  for (uint64_t mb : sizes_mb) {
    const uint64_t n_lines = (mb << 20) / 128;
    for (int bps : {4, 8}) {
      const int grid = prop.multiProcessorCount * bps;
      const int rounds = 64;
      for (int inflight : {4, 8}) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
          cudaEventRecord(e0);
          if (inflight == 4) probe<4><<<grid, 256>>>(buf, n_lines, rounds, 12345 + rep, sink);
          else probe<8><<<grid, 256>>>(buf, n_lines, rounds, 12345 + rep, sink);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        const double n = (double)grid * 256 * rounds * inflight;
        printf("footprint %6llu MiB  in-flight %d  blocks/SM %d : %.3f ms  %.3e lines/s  %.0f GB/s\n", (unsigned long long)mb, inflight, bps, best,
               n / (best * 1e-3), n * 128 / (best * 1e-3) / 1e9);
      }
    }
  }
  return 0;
}
