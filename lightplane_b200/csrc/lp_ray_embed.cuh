// Host-side glue of the renderer MODULE, fused:
//
//  * ray encoding = Linear(harmonic_embedding(normalize(direction)))   (renderer_module.py:578-601,
//    ray_utils.py:181-212): one kernel forward, one backward (weight / bias gradients) instead of
//    normalize + pow + mul + sin + cat + two SIMT sgemm + a column reduction;
//  * background epilogue = features + exp(-NLT) * bg and alpha = 1 - exp(-NLT) | -NLT
//    (renderer_module.py:552-563): one kernel each way.
//
// Byte work: the forward writes the [n, out] encoding once (coalesced float4 rows staged through
// shared memory), the backward reads the [n, out] encoding gradient once and keeps the
// (in+1) x out weight / bias gradient in registers across a block's tiles.
#pragma once

#include "lp_common.cuh"
#include "lp_platform.cuh"

#define LP_EMB_TILE 256       // rays per tile = threads per block
#define LP_EMB_MAX_HARM 10    // 3 + 6 * 10 = 63 embedding columns
#define LP_EMB_MAX_ITEMS 4    // per-thread accumulator groups of the backward

// embedding row of one ray, layout of ray_utils.py:181-212:
//   for phase in (0, pi/2): for coordinate in xyz: for k < H: sin(d * 2^k + phase);  then d itself
LP_DEVICE void lp_embed_row(const float* __restrict__ dir, long long ray, int H, float* e) {
  float d[3] = {dir[3 * ray], dir[3 * ray + 1], dir[3 * ray + 2]};
  const float nrm = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-12f);  // F.normalize(eps=1e-12)
#pragma unroll
  for (int c = 0; c < 3; ++c) d[c] = d[c] / nrm;
  for (int p = 0; p < 2; ++p)
    for (int c = 0; c < 3; ++c) {
      float f = 1.f;
      for (int k = 0; k < H; ++k, f *= 2.f)
        e[(p * 3 + c) * H + k] = sinf(d[c] * f + (p ? 1.5707963267948966f : 0.f));
    }
#pragma unroll
  for (int c = 0; c < 3; ++c) e[6 * H + c] = d[c];
}

// smem: W [out][in] | bias [out] | E [TILE][in + 1]
__global__ void lp_ray_embed_fwd_kernel(const float* __restrict__ dir, long long n, int H,
                                        const float* __restrict__ W, const float* __restrict__ bias, int out_dim,
                                        float* __restrict__ enc) {
  LP_DYN_SMEM(float, smem);
  const int in_dim = 3 + 6 * H, es = in_dim + 1, c4 = out_dim >> 2;
  float* sW = smem;
  float* sB = sW + out_dim * in_dim;
  float* sE = sB + out_dim;
  for (int i = threadIdx.x; i < out_dim * in_dim; i += blockDim.x) sW[i] = W[i];
  for (int i = threadIdx.x; i < out_dim; i += blockDim.x) sB[i] = bias ? bias[i] : 0.f;
  const long long tiles = (n + LP_EMB_TILE - 1) / LP_EMB_TILE;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    __syncthreads();
    const long long r0 = tile * LP_EMB_TILE;
    if (r0 + threadIdx.x < n) lp_embed_row(dir, r0 + threadIdx.x, H, sE + threadIdx.x * es);
    __syncthreads();
    for (int item = threadIdx.x; item < LP_EMB_TILE * c4; item += blockDim.x) {
      const int r = item / c4, q = item - r * c4;
      if (r0 + r >= n) break;
      const float* e = sE + r * es;
      const float* w = sW + 4 * q * in_dim;
      float4 a = make_float4(sB[4 * q], sB[4 * q + 1], sB[4 * q + 2], sB[4 * q + 3]);
      for (int i = 0; i < in_dim; ++i) {
        const float ev = e[i];
        a.x = fmaf(ev, w[i], a.x);
        a.y = fmaf(ev, w[in_dim + i], a.y);
        a.z = fmaf(ev, w[2 * in_dim + i], a.z);
        a.w = fmaf(ev, w[3 * in_dim + i], a.w);
      }
      *reinterpret_cast<float4*>(enc + (r0 + r) * out_dim + 4 * q) = a;
    }
  }
}

// smem: E [TILE][in + 2] (column `in` = 1 for the bias gradient) | G [TILE][out + 4]
// Thread item p = (embedding column i <= in, float4 chunk q of the outputs): 4 accumulators.
__global__ void lp_ray_embed_bwd_kernel(const float* __restrict__ dir, long long n, int H,
                                        const float* __restrict__ g_enc, int out_dim,
                                        float* __restrict__ g_W, float* __restrict__ g_bias) {
  LP_DYN_SMEM(float, smem);
  const int in_dim = 3 + 6 * H, es = in_dim + 2, gs = out_dim + 4, c4 = out_dim >> 2;
  float* sE = smem;
  float* sG = sE + LP_EMB_TILE * es;
  const int items = (in_dim + 1) * c4;
  float4 acc[LP_EMB_MAX_ITEMS];
#pragma unroll
  for (int a = 0; a < LP_EMB_MAX_ITEMS; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long tiles = (n + LP_EMB_TILE - 1) / LP_EMB_TILE;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    __syncthreads();
    const long long r0 = tile * LP_EMB_TILE;
    {
      float* e = sE + threadIdx.x * es;
      if (r0 + threadIdx.x < n) {
        lp_embed_row(dir, r0 + threadIdx.x, H, e);
        e[in_dim] = 1.f;
      } else {
        for (int i = 0; i <= in_dim; ++i) e[i] = 0.f;
      }
    }
    for (int item = threadIdx.x; item < LP_EMB_TILE * c4; item += blockDim.x) {
      const int r = item / c4, q = item - r * c4;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < n) g = lp_ldg4(g_enc + (r0 + r) * out_dim + 4 * q);
      *reinterpret_cast<float4*>(sG + r * gs + 4 * q) = g;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < LP_EMB_MAX_ITEMS; ++a) {
      const int p = threadIdx.x + a * LP_EMB_TILE;
      if (p < items) {
        const int i = p / c4, q = p - i * c4;
        const float* e = sE + i;
        const float* g = sG + 4 * q;
        float4 s = acc[a];
#pragma unroll 4
        for (int r = 0; r < LP_EMB_TILE; ++r) {
          const float ev = e[r * es];
          const float4 gv = *reinterpret_cast<const float4*>(g + r * gs);
          s.x = fmaf(ev, gv.x, s.x); s.y = fmaf(ev, gv.y, s.y);
          s.z = fmaf(ev, gv.z, s.z); s.w = fmaf(ev, gv.w, s.w);
        }
        acc[a] = s;
      }
    }
  }
#pragma unroll
  for (int a = 0; a < LP_EMB_MAX_ITEMS; ++a) {
    const int p = threadIdx.x + a * LP_EMB_TILE;
    if (p < items) {
      const int i = p / c4, q = p - i * c4;
      const float v[4] = {acc[a].x, acc[a].y, acc[a].z, acc[a].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i < in_dim) lp_red_add1(g_W + (4 * q + j) * in_dim + i, v[j]);
        else if (g_bias != nullptr) lp_red_add1(g_bias + 4 * q + j, v[j]);
      }
    }
  }
}

// ---- background epilogue -------------------------------------------------------------------
// alpha = log_t ? -nlt : 1 - exp(-nlt);  out[r, c] = feat[r, c] + exp(-nlt) * bg[c]
__global__ void lp_bg_fwd_kernel(long long n, int C, const float* __restrict__ nlt, const float* __restrict__ feat,
                                 const float* __restrict__ bg, int log_t, float* __restrict__ alpha,
                                 float* __restrict__ out) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const float l = nlt[r], T = expf(-l);
    alpha[r] = log_t ? -l : 1.f - T;
    for (int c = 0; c < C; ++c) out[r * C + c] = feat[r * C + c] + T * bg[c];
  }
}
// g_nlt[r] = (log_t ? -g_alpha : T * g_alpha) - T * sum_c g_out[r, c] * bg[c]     (d features = g_out, passed through by the caller)
__global__ void lp_bg_bwd_kernel(long long n, int C, const float* __restrict__ nlt, const float* __restrict__ bg,
                                 int log_t, const float* __restrict__ g_alpha, const float* __restrict__ g_out,
                                 float* __restrict__ g_nlt) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const float T = expf(-nlt[r]);
    float s = 0.f;
    if (g_out != nullptr)
      for (int c = 0; c < C; ++c) s = fmaf(g_out[r * C + c], bg[c], s);
    const float ga = g_alpha != nullptr ? g_alpha[r] : 0.f;
    g_nlt[r] = (log_t ? -ga : T * ga) - T * s;
  }
}
