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
// warp-level tensor-core primitives (mma.sync m16n8k8, TF32 in / FP32 accumulate)
// Fragment layout (lane = 4*g + t): A 16x8 row-major: a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4);
// B 8x8: b0=(k=t,n=g) b1=(k=t+4,n=g); C 16x8: c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1).
// ---------------------------------------------------------------------------------------------
LP_DEVICE float lp_tf32_rna(float x) {  // round-to-nearest TF32 (10-bit mantissa), as fp32 bits
#if defined(LP_HOSTSIM)
  unsigned u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0x00001000u;  // round half away from zero on the magnitude, like cvt.rna
  u &= 0xffffe000u;
  return __uint_as_float(u);
#else
  unsigned u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
#endif
}

LP_DEVICE void lp_mma_tf32(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
#if defined(LP_HOSTSIM)
  lp_hostsim_mma_m16n8k8(d, a, b);
#else
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
        "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
#endif
}
