// Validates and times the tcgen05 pieces of the thread-per-sample MLP chain, in isolation:
//   * tcgen05.st (registers -> TMEM) producing the A operand, tcgen05.mma kind::tf32 with A in TMEM and
//     B in shared memory (K-major, no swizzle), M=128 N=32 K=8 x 4 k-steps, commit -> mbarrier, tcgen05.ld;
//   * which of the two descriptor offset fields is the K-direction stride for a K-major operand;
//   * cycles: one st -> barrier -> 12 MMAs -> commit -> wait -> ld round trip; back-to-back MMA issue rate;
//     tcgen05.ld / tcgen05.st throughput with 4 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_test2 tc_test2.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t f1_bytes, uint32_t f2_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((f1_bytes >> 4) & 0x3FFF) << 16;  // "leading dimension byte offset" field
  d |= (uint64_t)((f2_bytes >> 4) & 0x3FFF) << 32;  // "stride dimension byte offset" field
  d |= (uint64_t)1 << 46;
  return d;
}
#define LD32(v, taddr)                                                                                                  \
  asm volatile(                                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                         \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, " \
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                                 \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),     \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),          \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),         \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                       \
      : "r"(taddr))
#define ST32(v, taddr)                                                                                                  \
  asm volatile(                                                                                                         \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "                                                                  \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, " \
      "%24, %25, %26, %27, %28, %29, %30, %31};" ::"r"(v[0]),                                                           \
      "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),     \
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),       \
      "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),       \
      "r"(v[29]), "r"(v[30]), "r"(v[31]), "r"(taddr)                                                                    \
      : "memory")

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int M = 128, N = 32, K = 32;
// K-major no-swizzle fp32 operand [rows][K]: core matrix = 8 rows x 16 B (4 elements of K), stored [row/8][k/4][8][4]
__host__ __device__ inline int kmajor_index(int row, int k, int kdim) { return ((row / 8) * (kdim / 4) + k / 4) * 32 + (row % 8) * 4 + (k % 4); }

// out[0..M*N): D = A * B^T ; times[0] = round trip cycles, [1] = cycles per MMA back to back, [2] = ld x32 per warp, [3] = st x32
__global__ void __launch_bounds__(128) tc2_kernel(const float* gA, const float* gB, float* out, long long* times, int variant) {
  __shared__ __align__(128) float sB[256 * K];  // first N*K used for numerics; the rest only for the N sweep
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N * K; i += blockDim.x) sB[i] = gB[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t A_COL = 64, D_COL = 0;
  // idesc: D=F32, A=B=TF32, K-major both, N, M
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
  const uint32_t kstride = 128, nstride = (K / 4) * 128;
  uint32_t a[32];
  for (int k = 0; k < 32; ++k) a[k] = __float_as_uint(gA[tid * K + k]);  // this thread's row of A

  long long t0 = clock64();
  ST32(a, tmem + lane_base + A_COL);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int rep = 0; rep < 3; ++rep)  // 12 MMAs like a 3xTF32 32x32 layer; reps 1,2 recompute the same product
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t db = variant == 0 ? make_desc(smem_u32(sB) + ks * 2 * kstride, kstride, nstride)
                                         : make_desc(smem_u32(sB) + ks * 2 * kstride, nstride, kstride);
        mma_tf32_ts(tmem + D_COL, tmem + A_COL + ks * 8, db, idesc, ks > 0);
      }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  LD32(v, tmem + lane_base + D_COL);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  long long t1 = clock64();
  for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(v[j]);
  if (tid == 0) times[0] = t1 - t0;

  // ---- MMA issue rate: 256 MMAs back to back (unrolled, descriptors precomputed), for several N ----
  uint32_t phase = 1;
  for (int ni = 0; ni < 5; ++ni) {
    const int n = 16 << ni;  // 16, 32, 64, 128, 256
    const uint32_t idn = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((M >> 4) << 24);
    __syncthreads();
    t0 = clock64();
    if (tid == 0) {
      const uint64_t db = make_desc(smem_u32(sB), kstride, nstride);
#pragma unroll 16
      for (int i = 0; i < 256; ++i) mma_tf32_ts(tmem + 256, tmem + A_COL + (i & 3) * 8, db, idn, 1);
      commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    t1 = clock64();
    if (tid == 0) times[4 + ni] = (t1 - t0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0) times[1] = times[5] / 256;
  // ---- ld / st throughput: 64 x (32 columns) per warp, all 4 warps ----
  __syncthreads();
  t0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 64; ++i) {
    LD32(v, tmem + lane_base + (i & 3) * 32);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += v[i & 31];
  }
  t1 = clock64();
  if (tid == 0) times[2] = (t1 - t0) / 64;
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < 64; ++i) {
    a[0] = acc + i;
    ST32(a, tmem + lane_base + A_COL + (i & 1) * 32);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  t1 = clock64();
  if (tid == 0) times[3] = (t1 - t0) / 64;
  if (acc == 0x12345678u) out[0] = 0.f;


  // ---- several issuing threads: lane 0 of warps 0..NI-1 each issue 256 MMAs (N=32) into its own accumulator ----
  for (int NI = 1; NI <= 4; NI *= 2) {
    __syncthreads();
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(NI)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    t0 = clock64();
    if ((tid & 31) == 0 && warp < NI) {
      const uint64_t db = make_desc(smem_u32(sB), kstride, nstride);
#pragma unroll 16
      for (int i = 0; i < 256; ++i) mma_tf32_ts(tmem + 256 + 32 * warp, tmem + A_COL + (i & 3) * 8, db, idesc, 1);
      commit(&bar2);
    }
    mbar_wait(&bar2, 0);
    t1 = clock64();
    if (tid == 0) times[9 + (NI == 1 ? 0 : NI == 2 ? 1 : 2)] = t1 - t0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  std::vector<float> A(M * K), W(K * N);  // D[m][n] = sum_k A[m][k] * W[k][n]
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m * K + k] = (float)(((m * 7 + k * 3) % 11) - 5) * 0.25f;
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) W[k * N + n] = (float)(((k * 5 + n * 2) % 7) - 3) * 0.5f;
  std::vector<float> hB(N * K);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[kmajor_index(n, k, K)] = W[k * N + n];
  float *dA, *dB, *dout; long long* dt;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dout, M * N * 4); cudaMalloc(&dt, 16 * 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dout, 0, M * N * 4);
    for (int rep = 0; rep < 2; ++rep) tc2_kernel<<<1, 128>>>(dA, dB, dout, dt, variant);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(M * N); long long t[16] = {0};
    cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(t, dt, 128, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * W[k * N + n];
      double err = fabs(ref - out[m * N + n]); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
    }
    printf("variant %d (%s): cuda=%s max_err=%g mismatches=%d/%d | cycles: round-trip(st,12 mma,ld)=%lld  per-mma=%lld  ld.x32=%lld  st.x32=%lld\n",
           variant, variant == 0 ? "field1=K stride, field2=N stride" : "swapped", cudaGetErrorString(e), maxerr, bad, M * N,
           t[0], t[1], t[2], t[3]);
    printf("   N=32, 256 MMAs per issuing thread, total cycles: 1 issuer %lld  2 issuers %lld  4 issuers %lld\n", t[9], t[10], t[11]);
    printf("   256 back-to-back tf32 MMAs (M=128,K=8), cycles per MMA: N=16 %.1f  N=32 %.1f  N=64 %.1f  N=128 %.1f  N=256 %.1f\n",
           t[4] / 256.0, t[5] / 256.0, t[6] / 256.0, t[7] / 256.0, t[8] / 256.0);
    if (e != cudaSuccess) break;
  }
  return 0;
}
