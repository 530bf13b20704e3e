// Validates the operand forms of the thread-per-sample forward chain:
//   tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), A in TMEM as packed bf16 pairs written by tcgen05.st,
//   B in shared memory K-major no-swizzle bf16; M=128, N=64, K=32; and the 2-term bf16 split
//   x = hi + lo with the three products hi*Whi + lo*Whi + hi*Wlo against an fp64 reference.
//   Second phase: the same weight tile read as an MN-major B operand (b_major = 1, field 1 = n-chunk stride of the
//   K-major tile, field 2 = 128) gives the transposed product A2 x W^T without a transposed copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_test3 tc_test3.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t kstride, uint32_t nstride) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((kstride >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((nstride >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db),
               "r"(idesc), "r"(acc) : "memory");
}
constexpr int M = 128, N = 64, K = 32;
__host__ __device__ inline int kmajor_bf16(int row, int k, int kdim) { return ((row / 8) * (kdim / 8) + k / 8) * 64 + (row % 8) * 8 + (k % 8); }

__device__ __forceinline__ uint32_t pack_hi(float a, float b) {  // truncated bf16 of a (low half) and b (high half)
  return (__float_as_uint(a) >> 16) | (__float_as_uint(b) & 0xFFFF0000u);
}
__device__ __forceinline__ float trunc_bf16(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_rn(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(128) tc3_kernel(const float* gA, const __nv_bfloat16* gBhi, const __nv_bfloat16* gBlo, float* out,
                                                  const float* gA2, float* out2) {
  __shared__ __align__(128) __nv_bfloat16 sBhi[N * K], sBlo[N * K];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N * K; i += blockDim.x) { sBhi[i] = gBhi[i]; sBlo[i] = gBlo[i]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t D_COL = 0, AH_COL = 64, AL_COL = 80;  // K=32 bf16 = 16 columns each
  // this thread's row of A: hi = truncated bf16, lo = rn bf16 of the exact remainder
  uint32_t ah[16], al[16];
  for (int j = 0; j < 16; ++j) {
    const float x0 = gA[tid * K + 2 * j], x1 = gA[tid * K + 2 * j + 1];
    ah[j] = pack_hi(x0, x1);
    al[j] = pack_rn(x0 - trunc_bf16(x0), x1 - trunc_bf16(x1));
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
               ::"r"(ah[0]), "r"(ah[1]), "r"(ah[2]), "r"(ah[3]), "r"(ah[4]), "r"(ah[5]), "r"(ah[6]), "r"(ah[7]), "r"(ah[8]), "r"(ah[9]),
               "r"(ah[10]), "r"(ah[11]), "r"(ah[12]), "r"(ah[13]), "r"(ah[14]), "r"(ah[15]), "r"(tmem + lane_base + AH_COL) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
               ::"r"(al[0]), "r"(al[1]), "r"(al[2]), "r"(al[3]), "r"(al[4]), "r"(al[5]), "r"(al[6]), "r"(al[7]), "r"(al[8]), "r"(al[9]),
               "r"(al[10]), "r"(al[11]), "r"(al[12]), "r"(al[13]), "r"(al[14]), "r"(al[15]), "r"(tmem + lane_base + AL_COL) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // idesc: D=F32 (1<<4), A=B=BF16 (1<<7, 1<<10), K-major both, N, M
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
    const uint32_t kstride = 128, nstride = (K / 8) * 128;
    int first = 1;
    for (int prod = 0; prod < 3; ++prod)   // hi*Whi, lo*Whi, hi*Wlo
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t a_col = (prod == 1 ? AL_COL : AH_COL) + ks * 8;
        const __nv_bfloat16* b = prod == 2 ? sBlo : sBhi;
        mma_bf16_ts(tmem + D_COL, tmem + a_col, make_desc(smem_u32(b) + ks * 2 * kstride, kstride, nstride), idesc, !first);
        first = 0;
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tmem + lane_base + D_COL + 32 * half));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[tid * N + 32 * half + j] = __uint_as_float(v[j]);
  }
  // ---- phase 2: D2[m][i] = sum_o A2[m][o] * W[i][o], i < K (= 32), o < N (= 64): hi x hi only ----
  {
    uint32_t a2[32];
    for (int j = 0; j < 32; ++j) a2[j] = pack_rn(gA2[tid * N + 2 * j], gA2[tid * N + 2 * j + 1]);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
                 ::"r"(a2[0]), "r"(a2[1]), "r"(a2[2]), "r"(a2[3]), "r"(a2[4]), "r"(a2[5]), "r"(a2[6]), "r"(a2[7]), "r"(a2[8]), "r"(a2[9]),
                 "r"(a2[10]), "r"(a2[11]), "r"(a2[12]), "r"(a2[13]), "r"(a2[14]), "r"(a2[15]), "r"(tmem + lane_base + AH_COL) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
                 ::"r"(a2[16]), "r"(a2[17]), "r"(a2[18]), "r"(a2[19]), "r"(a2[20]), "r"(a2[21]), "r"(a2[22]), "r"(a2[23]), "r"(a2[24]),
                 "r"(a2[25]), "r"(a2[26]), "r"(a2[27]), "r"(a2[28]), "r"(a2[29]), "r"(a2[30]), "r"(a2[31]), "r"(tmem + lane_base + AH_COL + 16) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // D = F32, A = B = BF16, B MN-major (bit 16), N = K (32), M = 128
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((K >> 3) << 17) | ((M >> 4) << 24);
      const uint32_t nstride = (K / 8) * 128;  // stride between 8-row (out) chunks of the K-major tile
      for (int ks = 0; ks < N / 16; ++ks)  // K' = N = 64 outputs, 16 per MMA = two chunks of 8 outs
        mma_bf16_ts(tmem + D_COL, tmem + AH_COL + ks * 8, make_desc(smem_u32(sBhi) + ks * 2 * nstride, nstride, 128), idesc2, ks > 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tmem + lane_base + D_COL));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out2[tid * K + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

int main() {
  std::vector<float> A(M * K), W(K * N);
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) / 16777216.0f) * 2.f - 1.f; };
  for (auto& v : A) v = rnd() * 3.f;
  for (auto& v : W) v = rnd();
  std::vector<__nv_bfloat16> hi(N * K), lo(N * K);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
    const float w = W[k * N + n];
    const __nv_bfloat16 h = __float2bfloat16(w);
    hi[kmajor_bf16(n, k, K)] = h;
    lo[kmajor_bf16(n, k, K)] = __float2bfloat16(w - __bfloat162float(h));
  }
  float *dA, *dout; __nv_bfloat16 *dhi, *dlo;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dhi, hi.size() * 2); cudaMalloc(&dlo, lo.size() * 2); cudaMalloc(&dout, M * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dhi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dlo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, M * N * 4);
  std::vector<float> A2(M * N);
  for (auto& v : A2) v = __bfloat162float(__float2bfloat16(rnd()));  // exactly representable
  float *dA2, *dout2;
  cudaMalloc(&dA2, A2.size() * 4); cudaMalloc(&dout2, M * K * 4);
  cudaMemcpy(dA2, A2.data(), A2.size() * 4, cudaMemcpyHostToDevice);
  tc3_kernel<<<1, 128>>>(dA, dhi, dlo, dout, dA2, dout2);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> out(M * N);
  cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
  double maxrel = 0, sumabs = 0, sumerr = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0, mag = 0;
    for (int k = 0; k < K; ++k) { ref += (double)A[m * K + k] * W[k * N + n]; mag += fabs((double)A[m * K + k] * W[k * N + n]); }
    const double err = fabs(ref - out[m * N + n]);
    sumerr += err; sumabs += fabs(ref);
    if (err / mag > maxrel) maxrel = err / mag;
  }
  printf("bf16 TMEM-A x K-major smem-B, 3-product split: cuda=%s  mean|err|/mean|ref|=%.3g  max err/sum|terms|=%.3g  out[0..2]=%g %g %g\n",
         cudaGetErrorString(e), sumerr / sumabs, maxrel, out[0], out[1], out[2]);
  std::vector<float> out2(M * K);
  cudaMemcpy(out2.data(), dout2, M * K * 4, cudaMemcpyDeviceToHost);
  double e2 = 0, r2 = 0;
  for (int m = 0; m < M; ++m) for (int i = 0; i < K; ++i) {
    double ref = 0;
    for (int o = 0; o < N; ++o) ref += (double)A2[m * N + o] * __bfloat162float(hi[kmajor_bf16(o, i, K)]);
    e2 += fabs(ref - out2[m * K + i]); r2 += fabs(ref);
  }
  printf("same tile as MN-major B (transposed product, hi x hi): mean|err|/mean|ref|=%.3g  out2[0..2]=%g %g %g\n", e2 / r2, out2[0], out2[1], out2[2]);
  return 0;
}
