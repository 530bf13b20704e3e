// Validates the tcgen05 pieces the backward kernel relies on, in isolation:
//   * TMEM alloc / dealloc, tcgen05.mma kind::f16 (bf16 in, fp32 accumulate) with BOTH operands
//     MN-major, no swizzle, M=128 N=32 K=16, accumulating two k-steps (32 "samples");
//   * descriptor encodings (LBO/SBO semantics for MN-major), commit -> mbarrier, tcgen05.ld.
// D[m][n] = sum_s A[m][s] * B[s][n]  with A = "X^T" (features x samples), B = "dY" (samples x out).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_test tc_test.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

// element (mn, k) of an MN-major no-swizzle bf16 operand: 16-byte chunks of 8 MN elements,
// 8 k-rows of 16 B form a core matrix; k-groups at LBO, MN-chunks at SBO
__host__ __device__ inline int mn_major_index(int mn, int k, int lbo_elems, int sbo_elems) {
  return (mn / 8) * sbo_elems + (k / 8) * lbo_elems + (k % 8) * 8 + (mn % 8);
}

constexpr int M = 128, N = 32, KTOT = 32;
constexpr int LBO = 128, SBO_A = 512, SBO_B = 512;  // bytes

__global__ void tc_kernel(const __nv_bfloat16* gA, const __nv_bfloat16* gB, float* out, int variant) {
  __shared__ __align__(128) __nv_bfloat16 sA[M * KTOT];
  __shared__ __align__(128) __nv_bfloat16 sB[N * KTOT];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * KTOT; i += blockDim.x) sA[i] = gA[i];
  for (int i = tid; i < N * KTOT; i += blockDim.x) sB[i] = gB[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy smem writes -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    // instruction descriptor: D=F32, A=B=BF16, both MN-major, N=32, M=128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
    for (int ks = 0; ks < KTOT / 16; ++ks) {
      uint32_t lbo = LBO, sbo_a = SBO_A, sbo_b = SBO_B;
      if (variant == 1) { uint32_t t = lbo; lbo = sbo_a; sbo_a = t; sbo_b = t; }  // swapped semantics probe
      const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO, lbo, sbo_a);
      const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * LBO, lbo, sbo_b);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  // everyone waits for the MMA to finish (phase 0)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // thread i of warp w reads TMEM lane 32w+i, 32 columns
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
  std::vector<float> A(M * KTOT), B(KTOT * N);
  for (int m = 0; m < M; ++m) for (int k = 0; k < KTOT; ++k) A[m * KTOT + k] = (float)(((m * 7 + k * 3) % 11) - 5);
  for (int k = 0; k < KTOT; ++k) for (int n = 0; n < N; ++n) B[k * N + n] = (float)(((k * 5 + n * 2) % 7) - 3);
  std::vector<__nv_bfloat16> hA(M * KTOT), hB(N * KTOT);
  for (int m = 0; m < M; ++m) for (int k = 0; k < KTOT; ++k) hA[mn_major_index(m, k, LBO / 2, SBO_A / 2)] = __float2bfloat16(A[m * KTOT + k]);
  for (int n = 0; n < N; ++n) for (int k = 0; k < KTOT; ++k) hB[mn_major_index(n, k, LBO / 2, SBO_B / 2)] = __float2bfloat16(B[k * N + n]);
  __nv_bfloat16 *dA, *dB; float* dout;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dout, M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dout, 0, M * N * 4);
    tc_kernel<<<1, 128>>>(dA, dB, dout, variant);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(M * N);
    cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0; for (int k = 0; k < KTOT; ++k) ref += (double)A[m * KTOT + k] * B[k * N + n];
      double err = fabs(ref - out[m * N + n]); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
    }
    printf("variant %d (%s): cuda=%s max_err=%g mismatches=%d/%d  out[0..3]=%g %g %g %g\n", variant,
           variant == 0 ? "LBO=k-group stride, SBO=MN-chunk stride" : "swapped", cudaGetErrorString(e), maxerr, bad, M * N,
           out[0], out[1], out[2], out[3]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
