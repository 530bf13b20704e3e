// Platform layer: the kernels are plain CUDA C++ for sm_100a.  When LP_HOSTSIM is defined (only by
// tests/hostsim/, never by the product build) the same sources compile with g++ against an
// SIMT emulation shim so that kernel logic can be debugged in the GPU-less build container.
#pragma once

#include <stdint.h>

#ifdef LP_HOSTSIM
#include "lp_hostsim.h"  // tests/hostsim/lp_hostsim.h (test infrastructure)
#define LP_IS_DEVICE_BUILD 0
#else
#include <cuda_runtime.h>
#define LP_IS_DEVICE_BUILD 1
#define LP_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; \
  type* name = reinterpret_cast<type*>(name##_raw)
#define LP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#define LP_DEVICE __device__ __forceinline__
#define LP_WARP 32
#define LP_FULL_MASK 0xffffffffu

// Vector atomic add of 4 consecutive floats (16-byte aligned).  sm_90+ has a native
// `red.global.add.v4.f32`; exposed by CUDA 12.x as atomicAdd(float4*, float4).
LP_DEVICE void lp_red_add4(float* addr, float a, float b, float c, float d) {
#if defined(LP_HOSTSIM)
  atomicAdd(addr + 0, a); atomicAdd(addr + 1, b); atomicAdd(addr + 2, c); atomicAdd(addr + 3, d);
#else
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
#endif
}

// predicated form (no branch): the reduction is issued only when `pred` holds
LP_DEVICE void lp_red_add4_if(bool pred, float* addr, float a, float b, float c, float d) {
#if defined(LP_HOSTSIM)
  if (pred) { atomicAdd(addr + 0, a); atomicAdd(addr + 1, b); atomicAdd(addr + 2, c); atomicAdd(addr + 3, d); }
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}\n" ::"l"(addr),
      "f"(a), "f"(b), "f"(c), "f"(d), "r"((int)pred)
      : "memory");
#endif
}

LP_DEVICE void lp_red_add1(float* addr, float a) {
#if defined(LP_HOSTSIM)
  atomicAdd(addr, a);
#else
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
#endif
}

LP_DEVICE float4 lp_ldg4(const float* p) {
#if defined(LP_HOSTSIM)
  return make_float4(p[0], p[1], p[2], p[3]);
#else
  return __ldg(reinterpret_cast<const float4*>(p));
#endif
}

// ---------------------------------------------------------------------------------------------
// Blackwell tensor-core primitives (tcgen05): accumulators and the A operand of the per-sample MLP
// chain live in tensor memory, B operands and the parameter-gradient tiles in shared memory as
// non-swizzled UMMA operands.  The operand forms and descriptor encodings used below are validated
// in isolation by tools/tc_test.cu, tc_test2.cu and tc_test3.cu.
// ---------------------------------------------------------------------------------------------
LP_DEVICE unsigned lp_pack_bf16x2(float lo, float hi) {  // lo -> bits 0..15, hi -> bits 16..31
#if defined(LP_HOSTSIM)
  auto cv = [](float x) -> unsigned {
    unsigned u = __float_as_uint(x);
    if ((u & 0x7f800000u) == 0x7f800000u) return u >> 16;
    u += 0x7fffu + ((u >> 16) & 1u);  // round to nearest even
    return u >> 16;
  };
  return cv(lo) | (cv(hi) << 16);
#else
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
#endif
}

// ReLU fused into the conversion / into a packed pair: max(x, 0) rounded to bf16; max(bf16 pair, 0)
LP_DEVICE unsigned lp_pack_bf16x2_relu(float lo, float hi) {
#if defined(LP_HOSTSIM)
  return lp_pack_bf16x2(lo > 0.f ? lo : 0.f, hi > 0.f ? hi : 0.f);
#else
  unsigned r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
#endif
}
LP_DEVICE unsigned lp_relu_bf16x2(unsigned v) {
#if defined(LP_HOSTSIM)
  auto h = [](unsigned x) -> unsigned { return ((x & 0x8000u) || (x & 0x7fffu) > 0x7f80u) ? 0u : x; };  // negative or NaN -> +0
  return h(v & 0xffffu) | (h(v >> 16) << 16);
#else
  unsigned r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(0u));
  return r;
#endif
}

// fp32 -> tf32 (10-bit mantissa) with round-to-nearest, as a 32-bit pattern; the tensor core ignores the low 13 bits
// of a kind::tf32 operand, so feeding raw fp32 would truncate (a systematic bias of -2^-12 per factor)
LP_DEVICE unsigned lp_tf32_rna(float x) {
#if defined(LP_HOSTSIM)
  unsigned u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return u;
  return (u + 0x1000u) & 0xffffe000u;
#else
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
#endif
}
// ReLU gate from a packed pair of bf16 activations: a (b) survives iff the low (high) half of w is non-zero
LP_DEVICE void lp_gate2(unsigned w, float& a, float& b) {
#if defined(LP_HOSTSIM)
  if (!(w & 0x7fffu)) a = 0.f;
  if (!((w >> 16) & 0x7fffu)) b = 0.f;
#else
  asm("{\n\t.reg .pred p, q;\n\tsetp.ne.bf16x2 p|q, %2, %3;\n\tselp.f32 %0, %0, 0f00000000, p;\n\t"
      "selp.f32 %1, %1, 0f00000000, q;\n\t}\n" : "+f"(a), "+f"(b) : "r"(w), "r"(0u));
#endif
}
// approximate transcendentals of the tensor-core kernels' compositing (ex2.approx / lg2.approx / rcp.approx, relative
// error ~2^-22): the forward kernel and the backward kernel's recompute use the same ones, so the saved outputs and
// the recomputed prefix sums stay consistent
LP_DEVICE float lp_fast_exp(float x) {
#if defined(LP_HOSTSIM)
  return expf(x);
#else
  return __expf(x);
#endif
}
LP_DEVICE float lp_fast_rcp(float x) {
#if defined(LP_HOSTSIM)
  return 1.f / x;
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
LP_DEVICE float lp_fast_log(float x) {
#if defined(LP_HOSTSIM)
  return logf(x);
#else
  return __logf(x);
#endif
}

#if !defined(LP_HOSTSIM)
LP_DEVICE unsigned lp_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

LP_DEVICE void lp_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lp_smem_u32(bar)), "r"(count));
}
LP_DEVICE void lp_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
LP_DEVICE void lp_mbar_wait(unsigned long long* bar, int parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        // (suspend-time hint: the thread may sleep in hardware up to ~20 us before the test returns false; without it
        // waiting warps re-issued the test back to back -- 9 % of all instructions of the warp-specialised backward)
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(lp_smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
  }
}
// One lane of the (converged) warp: `elect.sync`.  Unlike `lane == 0` the compiler knows that exactly one thread passes,
// so single-thread instructions behind it (tcgen05.mma, tcgen05.commit) are emitted without an election loop.
LP_DEVICE bool lp_elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
// plain arrival of one thread (release semantics at CTA scope)
LP_DEVICE void lp_mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(lp_smem_u32(bar)) : "memory");
}
LP_DEVICE void lp_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
LP_DEVICE void lp_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
LP_DEVICE void lp_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// warp-collective (call from exactly one full warp)
LP_DEVICE void lp_tmem_alloc512(unsigned* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(lp_smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
LP_DEVICE void lp_tmem_dealloc512(unsigned base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
LP_DEVICE void lp_tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(lp_smem_u32(bar)) : "memory");
}

// ---- TMEM-sourced ("TS") MMAs of the thread-per-sample MLP chain (lp_render_tc.cuh): A lives in TMEM
// (row m of A in TMEM lane m, written by tcgen05.st), B in shared memory, K-major, no swizzle:
//   bf16:  element (n, k) at (n/8)*nstride + (k/8)*128 + (n%8)*16 + (k%8)*2, two A elements per column
//   tf32:  element (n, k) at (n/8)*nstride + (k/4)*128 + (n%8)*16 + (k%4)*4, one A element per column
// Descriptor field 1 (bits 16..29) is the K-direction core-matrix stride, field 2 (bits 32..45) the
// N-direction stride (validated by tools/tc_test2.cu / tc_test3.cu).  One MMA covers K = 32 bytes.
LP_DEVICE unsigned lp_tc_kdesc_lo(const void* smem_ptr) { return ((lp_smem_u32(smem_ptr) >> 4) & 0x3FFF) | (8u << 16); }
typedef unsigned lp_kdesc_t;
LP_DEVICE lp_kdesc_t lp_tc_kadv(lp_kdesc_t lo, int bytes) { return lo + (unsigned)(bytes >> 4); }
LP_DEVICE void lp_tc_mma_ts(bool tf32, unsigned d_taddr, unsigned a_taddr, unsigned b_lo, int nstride, int n, int accumulate) {
#ifdef LP_ABL_NO_MMA  // profiling only: hand-offs and waits stay, the tensor core does nothing
  return;
#endif
  const unsigned fmt = tf32 ? 2u : 1u;
  const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(n >> 3) << 17) | (8u << 24);
  const unsigned hi = (unsigned)(nstride >> 4) | (1u << 14);
  if (tf32)
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d_taddr), "r"(a_taddr), "r"(b_lo),
                 "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d_taddr), "r"(a_taddr), "r"(b_lo),
                 "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
// TS MMA whose B operand is a K-major weight tile read TRANSPOSED (as an MN-major operand, b_major = 1): for a
// tile [n][k] with n-chunk stride `nstride`, B'[n' = k][k' = n] sits at (k'/8)*nstride + (n'/8)*128 + (k'%8)*16 +
// (n'%8)*2, i.e. descriptor field 1 (K-group stride) = nstride, field 2 (MN-chunk stride) = 128; one MMA consumes 16
// values of k' = two n-chunks of the tile (advance the start address by 2*nstride).  tools/tc_test3.cu, phase 2.
LP_DEVICE lp_kdesc_t lp_tc_kdesc_lo_t(const void* smem_ptr, int nstride) {
  return ((lp_smem_u32(smem_ptr) >> 4) & 0x3FFF) | ((unsigned)(nstride >> 4) << 16);
}
LP_DEVICE void lp_tc_mma_ts_t(unsigned d_taddr, unsigned a_taddr, lp_kdesc_t b_lo, int nstride, int n, int accumulate) {
  const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((unsigned)(n >> 3) << 17) | (8u << 24);
  const unsigned hi = 8u | (1u << 14);
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d_taddr), "r"(a_taddr), "r"(b_lo),
               "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
// both operands in shared memory, bf16, MN-major, no swizzle, K-group stride 128 B, MN-chunk stride `sbo`:
// element (mn, k) at (mn/8)*sbo + (k/8)*128 + (k%8)*16 + (mn%8)*2.  D(128 x n) (+)= A(128 x 16) * B(16 x n).
LP_DEVICE lp_kdesc_t lp_tc_mndesc_lo(const void* smem_ptr) { return ((lp_smem_u32(smem_ptr) >> 4) & 0x3FFF) | (8u << 16); }
LP_DEVICE void lp_tc_mma_ss_mn(unsigned d_taddr, lp_kdesc_t a_lo, lp_kdesc_t b_lo, int sbo, int n, int accumulate) {
#ifdef LP_ABL_NO_MMA
  return;
#endif
  const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(n >> 3) << 17) | (8u << 24);
  const unsigned hi = (unsigned)(sbo >> 4) | (1u << 14);
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(d_taddr), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc),
      "r"(accumulate) : "memory");
}
// warp-collective: lane i writes NW consecutive columns of TMEM lane 32*(warp%4)+i
template <int NW>
LP_DEVICE void lp_tmem_st(unsigned taddr, const unsigned (&v)[NW]) {
  static_assert(NW == 4 || NW == 8 || NW == 16 || NW == 32, "4, 8, 16 or 32 columns");
  if constexpr (NW == 32) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31};" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
        "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]), "r"(taddr) : "memory");
  } else if constexpr (NW == 4)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], {%0, %1, %2, %3};" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(taddr) : "memory");
  else if constexpr (NW == 8)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::"r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(taddr) : "memory");
  else
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
                 ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                 "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr) : "memory");
}
// clear N columns of this thread's TMEM lane (accumulators are pre-zeroed so that every MMA accumulates
// and the MMAs of one product can be issued by several threads in any order)
template <int N>
LP_DEVICE void lp_tmem_zero(unsigned taddr) {
  unsigned z[N];
#pragma unroll
  for (int j = 0; j < N; ++j) z[j] = 0u;
  lp_tmem_st<N>(taddr, z);
}
LP_DEVICE void lp_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// this thread's TMEM address for column `col`: lanes 32*(warp%4).. of the CTA's allocation
LP_DEVICE unsigned lp_taddr(unsigned tmem_base, int warp_in_group, int col) {
  return tmem_base + ((unsigned)(warp_in_group * 32) << 16) + (unsigned)col;
}
LP_DEVICE void lp_tmem_ld32u(unsigned taddr, float (&v)[32]) {
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
// warp-collective: lane i receives N consecutive columns of its TMEM lane
template <int N>
LP_DEVICE void lp_tmem_ld(unsigned taddr, float (&v)[N]) {
  static_assert(N == 8 || N == 16 || N == 32, "8, 16 or 32 columns");
  if constexpr (N == 32) {
    lp_tmem_ld32u(taddr, v);
  } else if constexpr (N == 16) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  } else {
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
  }
}
// register re-allocation between warp-specialised roles (all warps of a 128-thread warpgroup execute it)
#define LP_SETMAXNREG_INC(n) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(n))
#define LP_SETMAXNREG_DEC(n) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(n))
// named barrier over `nthreads` threads (ids 1..15; 0 is __syncthreads)
LP_DEVICE void lp_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// named barrier with an OR reduction: true iff `pred` holds for any of the `nthreads` threads
LP_DEVICE bool lp_bar_any(int id, int nthreads, bool pred) {
  unsigned r;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %3, 0;\n\tbarrier.red.or.pred p, %1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(r) : "r"(id), "r"(nthreads), "r"((unsigned)pred) : "memory");
  return r != 0;
}
#endif  // !LP_HOSTSIM

// ---- packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2: `*.f32x2`): one issue slot for two IEEE fp32 operations, each
// rounded exactly like its scalar form (tools/ubench_f32x2.cu: an FFMA2 occupies the FMA pipe for two cycles but the
// scheduler for one -- the tensor-core kernels are bound by issue slots, not by the FMA pipe).
LP_DEVICE float2 lp_fma2(float2 a, float2 b, float2 c) {
#if defined(LP_HOSTSIM)
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#else
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
#endif
}
#if defined(LP_HOSTSIM)
LP_DEVICE float2 lp_add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
LP_DEVICE float2 lp_sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
LP_DEVICE float2 lp_mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#else
#define LP_F32X2_BINARY(NAME, OP)                                                                                         \
  LP_DEVICE float2 NAME(float2 a, float2 b) {                                                                             \
    float2 d;                                                                                                             \
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t" OP ".rn.f32x2 rd, ra, rb;\n\t"       \
        "mov.b64 {%0, %1}, rd;\n\t}" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                    \
    return d;                                                                                                             \
  }
LP_F32X2_BINARY(lp_add2, "add")
LP_F32X2_BINARY(lp_sub2, "sub")
LP_F32X2_BINARY(lp_mul2, "mul")
#undef LP_F32X2_BINARY
#endif
LP_DEVICE float2 lp_f2(float x, float y) { return make_float2(x, y); }

// Broadcast of lane 0's value: tells the compiler the value is warp-uniform (it may then live in uniform registers).
#define LP_WARP_UNIFORM(x) __shfl_sync(0xffffffffu, (x), 0)
