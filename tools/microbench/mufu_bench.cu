// Throughput of the exponential paths available to the attention softmax on sm_100a:
//   ex2.approx.ftz.f32, ex2.approx.ftz.bf16x2 (2 results per op), and an FMA-only polynomial exp2.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2bf2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  uint32_t u[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); u[i] = 0xbf80bf80u + threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.0f;
      else if (MODE == 1) u[i] = ex2bf2(u[i]) ^ 0x80008000u;
      else if (MODE == 2) u[i] = ex2h2(u[i]) ^ 0x80008000u;
      else {
        // polynomial exp2 on FMA pipe: x = n + f, 2^f by degree-3 poly, exponent add
        float x = a[i];
        float tt = x + 12582912.0f;
        float n = tt - 12582912.0f;
        float f = x - n;
        float p = fmaf(fmaf(fmaf(0.0555f, f, 0.2402f), f, 0.6931f), f, 1.0f);
        a[i] = __int_as_float(__float_as_int(p) + (__float_as_int(tt) << 23)) - 1.5f;
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, int per_op) {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(d, 64);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 8 * 256 * (double)iters * 8 * per_op;
  printf("%-28s %.3f ms  %.2f Texp/s  (%.1f exp/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e9, ops / ms / 1e6 / 148 / 1.9e3 * 1e0);
  cudaFree(d);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  run<3>("poly exp2 (FMA pipe)", 1);
  return 0;
}
