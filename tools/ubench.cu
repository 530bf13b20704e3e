// Micro-benchmarks that size the design decisions of the fast Renderer path on the actual B200:
// FP32 FMA rate, legacy mma.sync TF32/BF16 rate, vector-atomic (red.v4) and 16-byte gather
// throughput on an L2-resident 786 KB table (the 64^2x16 triplane) and on a 256 MiB table.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run: ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_ffma(float* out, int iters) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float b = 1.0001f, c = 0.5f;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void k_mma_tf32(float* out, int iters) {
  float c[4][4] = {};
  unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int t = 0; t < 4; ++t) for (int j = 0; j < 4; ++j) s += c[t][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma_bf16(float* out, int iters) {
  float c[4][4] = {};
  unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f003f00u};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int t = 0; t < 4; ++t) for (int j = 0; j < 4; ++j) s += c[t][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s; }

// each quad (4 lanes) hits one random 64-byte row; coherent=1: all quads of a warp hit the same row
template <int MODE>  // 0: red.v4, 1: red scalar x4, 2: ldg.128 gather
__global__ void k_table(float* table, unsigned rows, int iters, int coherent, float* out) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) / (coherent ? 32 : 4) * 2654435761u + 12345u;
  const int t = threadIdx.x & 3;
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    unsigned r = lcg(s) % rows;
    float* p = table + (size_t)r * 16 + 4 * t;
    if (MODE == 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    } else if (MODE == 1) {
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.f) : "memory");
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 1), "f"(1.f) : "memory");
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 2), "f"(1.f) : "memory");
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 3), "f"(1.f) : "memory");
    } else {
      float4 v = __ldg(reinterpret_cast<const float4*>(p));
      acc += v.x + v.y + v.z + v.w;
    }
  }
  if (MODE == 2) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
float time_ms(F f, int reps = 3) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, clock attr %d kHz\n", p.name, sms, clk_khz);
  float* out; CK(cudaMalloc(&out, sizeof(float) * 148 * 32 * 1024));
  {
    int iters = 20000, blocks = sms * 4, threads = 512;
    float ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, iters); });
    double fma = (double)blocks * threads * iters * 8;
    printf("FFMA: %.2f TFLOP/s fp32  (%.1f FMA/ns/SM)\n", 2 * fma / ms / 1e9, fma / ms / 1e6 / sms);
  }
  for (int warps : {4, 8, 16}) {
    int iters = 20000, blocks = sms, threads = warps * 32;
    float ms = time_ms([&] { k_mma_tf32<<<blocks, threads>>>(out, iters); });
    double n = (double)blocks * warps * iters * 4;
    printf("mma.sync m16n8k8 tf32, %2d warps/SM: %.1f TFLOP/s, %.2f mma/ns/SM\n", warps, n * 2048 / ms / 1e9, n / ms / 1e6 / sms);
    ms = time_ms([&] { k_mma_bf16<<<blocks, threads>>>(out, iters); });
    printf("mma.sync m16n8k16 bf16, %2d warps/SM: %.1f TFLOP/s, %.2f mma/ns/SM\n", warps, n * 4096 / ms / 1e9, n / ms / 1e6 / sms);
  }
  for (size_t rows : {(size_t)12288, (size_t)4 << 20}) {
    float* table; CK(cudaMalloc(&table, rows * 64)); CK(cudaMemset(table, 0, rows * 64));
    int iters = 2000, blocks = sms * 8, threads = 256;
    double quads = (double)blocks * threads / 4 * iters;
    for (int coh = 0; coh < 2; ++coh) {
      float ms = time_ms([&] { k_table<0><<<blocks, threads>>>(table, (unsigned)rows, iters, coh, out); });
      printf("table %7zu rows (%6.1f MB) coherent=%d  red.v4 : %.2f G rows(64B)/s = %.0f GB/s\n", rows, rows * 64 / 1e6, coh, quads / ms / 1e6, quads * 64 / ms / 1e6);
      ms = time_ms([&] { k_table<1><<<blocks, threads>>>(table, (unsigned)rows, iters, coh, out); });
      printf("table %7zu rows (%6.1f MB) coherent=%d  red.f32: %.2f G rows(64B)/s = %.0f GB/s\n", rows, rows * 64 / 1e6, coh, quads / ms / 1e6, quads * 64 / ms / 1e6);
      ms = time_ms([&] { k_table<2><<<blocks, threads>>>(table, (unsigned)rows, iters, coh, out); });
      printf("table %7zu rows (%6.1f MB) coherent=%d  ldg.128: %.2f G rows(64B)/s = %.0f GB/s\n", rows, rows * 64 / 1e6, coh, quads / ms / 1e6, quads * 64 / ms / 1e6);
    }
    cudaFree(table);
  }
  return 0;
}
