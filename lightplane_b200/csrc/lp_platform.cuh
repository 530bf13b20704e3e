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
