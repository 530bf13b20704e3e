// Issue rate of packed fp32 (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a, alone and mixed with
// integer work: does one FFMA2 take one issue slot?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE>  // 0: 8 FFMA / iter, 1: 8 FFMA2 / iter, 2: 8 FFMA + 8 IADD3-ish, 3: 8 FFMA2 + 8 int
__global__ void k(float* out, int iters) {
  float a[8]; u64 p[8]; unsigned q[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); q[i] = threadIdx.x * 7 + i; }
  const float b = 1.0001f, c = 0.5f;
  const u64 b2 = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b), c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2) a[i] = fma1(a[i], b, c); else p[i] = fma2(p[i], b2, c2);
      if (MODE >= 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(0x9e3779b9u));
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* d, int threads) {
  const int iters = 20000, blocks = 148;
  k<MODE><<<blocks, threads>>>(d, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma_instr = 8.0 * iters * (threads / 32) * blocks;  // warp-level FMA instructions
  const double clk = ms * 1e-3 * 1.965e9;
  printf("%-28s threads/SM %4d  %.3f ms  FMA warp-instr / clk / SM = %.2f  (fp32 FMA lanes / clk / SM = %.0f)\n", name, threads, ms,
         fma_instr / clk / blocks, fma_instr / clk / blocks * 32 * ((MODE & 1) ? 2 : 1));
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("FFMA", d, threads); run<1>("FFMA2", d, threads); run<2>("FFMA + LOP3", d, threads); run<3>("FFMA2 + LOP3", d, threads);
  }
  return 0;
}
