// Splatter kernels.
//
// Plain splatter (no MLP): HBM/L2-atomic bound byte work.  A sub-warp of `lpr` lanes owns one ray
// and each lane a float4 chunk of the channels, so every tap of every sample is one coalesced
// `red.global.add.v4.f32` row segment (forward) or one 16-byte gather per lane (backward); feature
// and weight grid are accumulated in ONE march (the reference launches its kernel twice,
// lightplane_splatter.py:505,539).  Semantics: splatter_fw.py:71-165, splatter_bw.py:75-180.
//
// MLP splatter: generic lane-per-ray kernels built from the same blocks as the generic renderer
// (splatter_fw.py:168-309, splatter_bw.py:183-394).
#pragma once

#include "lp_render_generic.cuh"

template <int VPL>
__global__ void lp_splat_fwd_kernel(LpRays R, LpMarch M, LpGridSet OUT, float* __restrict__ weight,
                                    const float* __restrict__ valid, int lpr) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % lpr, rays_per_warp = LP_WARP / lpr;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long ray = warp_global * rays_per_warp + lane / lpr;
  if (ray >= R.n) return;  // no block-level barriers below
  const int C = OUT.C;
  const float vm = valid ? valid[ray] : 1.f;
  float4 f[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    f[v] = lp_ldg4(R.enc + ray * C + 4 * (sub + lpr * v));
    f[v].x *= vm; f[v].y *= vm; f[v].z *= vm; f[v].w *= vm;
  }
  const float ox = R.org[3 * ray], oy = R.org[3 * ray + 1], oz = R.org[3 * ray + 2];
  const float dx = R.dir[3 * ray], dy = R.dir[3 * ray + 1], dz = R.dir[3 * ray + 2];
  const float near = R.near[ray], far = R.far[ray];
  const int b = min(max(R.gidx[ray], 0), OUT.g[0].B - 1);
  const int tot = M.S + M.S_inf;
  for (int step = 0; step < tot; ++step) {
    const float depth = lp_depth(step, near, far, M.S, M.S_inf, M.disparity_at_inf);
    float x = ox + depth * dx, y = oy + depth * dy, z = oz + depth * dz;
    if (M.contract) lp_contract(x, y, z);
    if (M.mask_oob && lp_in_bounds(x, y, z) == 0.f) continue;
    for (int gi = 0; gi < OUT.n; ++gi) {
      long long off[8];
      float w[8];
      const int nt = lp_taps(OUT.g[gi], C, b, x, y, z, off, w);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        if (t < nt && w[t] != 0.f) {
#pragma unroll
          for (int v = 0; v < VPL; ++v)
            lp_red_add4(OUT.data + off[t] + 4 * (sub + lpr * v), w[t] * f[v].x, w[t] * f[v].y,
                        w[t] * f[v].z, w[t] * f[v].w);
          if (sub == 0 && weight != nullptr && vm != 0.f) lp_red_add1(weight + off[t] / C, w[t] * vm);
        }
      }
    }
  }
}

template <int VPL>
__global__ void lp_splat_bwd_kernel(LpRays R, LpMarch M, LpGridSet GG, const float* __restrict__ valid,
                                    float* __restrict__ g_feat, int lpr) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % lpr, rays_per_warp = LP_WARP / lpr;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long ray = warp_global * rays_per_warp + lane / lpr;
  if (ray >= R.n) return;
  const int C = GG.C;
  const float vm = valid ? valid[ray] : 1.f;
  float4 acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float ox = R.org[3 * ray], oy = R.org[3 * ray + 1], oz = R.org[3 * ray + 2];
  const float dx = R.dir[3 * ray], dy = R.dir[3 * ray + 1], dz = R.dir[3 * ray + 2];
  const float near = R.near[ray], far = R.far[ray];
  const int b = min(max(R.gidx[ray], 0), GG.g[0].B - 1);
  const int tot = M.S + M.S_inf;
  for (int step = 0; step < tot; ++step) {
    const float depth = lp_depth(step, near, far, M.S, M.S_inf, M.disparity_at_inf);
    float x = ox + depth * dx, y = oy + depth * dy, z = oz + depth * dz;
    if (M.contract) lp_contract(x, y, z);
    if (M.mask_oob && lp_in_bounds(x, y, z) == 0.f) continue;
    for (int gi = 0; gi < GG.n; ++gi) {
      long long off[8];
      float w[8];
      const int nt = lp_taps(GG.g[gi], C, b, x, y, z, off, w);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        if (t < nt && w[t] != 0.f) {
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const float4 g = lp_ldg4(GG.data + off[t] + 4 * (sub + lpr * v));
            acc[v].x = fmaf(w[t], g.x, acc[v].x); acc[v].y = fmaf(w[t], g.y, acc[v].y);
            acc[v].z = fmaf(w[t], g.z, acc[v].z); acc[v].w = fmaf(w[t], g.w, acc[v].w);
          }
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    float* o = g_feat + ray * C + 4 * (sub + lpr * v);
    o[0] = acc[v].x * vm; o[1] = acc[v].y * vm; o[2] = acc[v].z * vm; o[3] = acc[v].w * vm;
  }
}

// feat[r,:] /= max(w[r],1e-5); w[r] = max(w[r],1e-5)   (lightplane_splatter.py:541,584)
__global__ void lp_splat_normalize_kernel(float* __restrict__ feat, float* __restrict__ weight,
                                          long long rows, int C) {
  const int c4 = C >> 2;
  const long long total = rows * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4;
    const float w = fmaxf(weight[r], 1e-5f);
    float4* p = reinterpret_cast<float4*>(feat) + i;
    float4 v = *p;
    const float inv = 1.f / w;
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    *p = v;
    if (i - r * c4 == 0) weight[r] = w;
  }
}

__global__ void lp_int_to_randn_kernel(const int* __restrict__ x1, const int* __restrict__ x2,
                                       int seed, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = lp_int_to_randn(x1[i], x2[i], seed);
}

// -------------------------------------------------------------------------------------------
// MLP splatter (generic, lane = ray)
// -------------------------------------------------------------------------------------------
struct LpSplatMlp {
  LpMlp mlp;
  int n_params;
  int c_in, c_out;
  int max_dim;
  int x0, xin;             // arena slots: sampled input feature, MLP input
  int y[LP_MAX_LAYERS];    // layer outputs
  int total;
};

LP_DEVICE void lp_lane_splat_weight(const LpGridSet& G, float* weight, int b, float x, float y,
                                    float z, float scale) {
  if (scale == 0.f || weight == nullptr) return;
  for (int gi = 0; gi < G.n; ++gi) {
    long long off[8];
    float w[8];
    const int nt = lp_taps(G.g[gi], G.C, b, x, y, z, off, w);
    for (int t = 0; t < nt; ++t)
      if (w[t] != 0.f) lp_red_add1(weight + off[t] / G.C, w[t] * scale);
  }
}

LP_DEVICE const float* lp_eval_splat_mlp(const LpSplatMlp& S, const float* P, float* arena,
                                         const float* feat, const LpGridSet& IN, int b, float x,
                                         float y, float z, float oob, int lane) {
  float* x0 = arena + S.x0 * LP_LS;
  lp_lane_sample(IN, b, x, y, z, oob, false, x0, lane);
  float* xin = arena + S.xin * LP_LS;
  for (int k = 0; k < S.c_in; ++k) xin[k * LP_LS + lane] = x0[k * LP_LS + lane] + feat[k * LP_LS + lane];
  const float* h = xin;
  for (int l = 0; l < S.mlp.n_layers; ++l) {
    float* out = arena + S.y[l] * LP_LS;
    lp_lane_linear(P, S.mlp.l[l], h, out, lane);
    h = out;
  }
  return h;
}

__global__ void lp_mlp_splat_fwd_kernel(LpRays R, LpMarch M, LpSplatMlp S, LpGridSet IN,
                                        LpGridSet OUT, float* __restrict__ weight,
                                        const float* __restrict__ valid,
                                        const float* __restrict__ params, int params_in_smem) {
  LP_DYN_SMEM(float, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int pfloats = params_in_smem ? ((S.n_params + 3) & ~3) : 0;
  const float* P = params;
  if (params_in_smem) {
    for (int i = threadIdx.x; i < S.n_params; i += blockDim.x) smem[i] = params[i];
    P = smem;
  }
  __syncthreads();
  const int per_warp = (S.total + S.c_in + S.c_out) * LP_LS;
  float* arena = smem + pfloats + warp * per_warp;
  float* feat = arena + S.total * LP_LS;
  float* tmp = feat + S.c_in * LP_LS;
  const int ray = (blockIdx.x * nwarps + warp) * LP_WARP + lane;
  const LpRayState s = lp_load_ray(R, ray, OUT.g[0].B);
  const int rr = s.active ? ray : R.n - 1;
  const float vm = s.active ? (valid ? valid[rr] : 1.f) : 0.f;
  for (int k = 0; k < S.c_in; ++k) feat[k * LP_LS + lane] = R.enc[(long long)rr * S.c_in + k];
  const int tot = M.S + M.S_inf;
  for (int step = 0; step < tot; ++step) {
    const float depth = lp_depth(step, s.near, s.far, M.S, M.S_inf, M.disparity_at_inf);
    float x = s.ox + depth * s.dx, y = s.oy + depth * s.dy, z = s.oz + depth * s.dz;
    if (M.contract) lp_contract(x, y, z);
    const float oob = M.mask_oob ? lp_in_bounds(x, y, z) : 1.f;
    const float* yout = lp_eval_splat_mlp(S, P, arena, feat, IN, s.b, x, y, z, oob, lane);
    for (int c = 0; c < S.c_out; ++c) tmp[c * LP_LS + lane] = yout[c * LP_LS + lane];
    lp_lane_splat(OUT, OUT.data, s.b, x, y, z, oob * vm, tmp, lane);
    lp_lane_splat_weight(OUT, weight, s.b, x, y, z, oob * vm);
  }
}

__global__ void lp_mlp_splat_bwd_kernel(LpRays R, LpMarch M, LpSplatMlp S, LpGridSet IN, LpGridSet GG,
                                        const float* __restrict__ valid,
                                        const float* __restrict__ params, int params_in_smem,
                                        float* __restrict__ g_feat, float* __restrict__ g_params,
                                        float* __restrict__ g_in) {
  LP_DYN_SMEM(float, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int pfloats = (S.n_params + 3) & ~3;
  const float* P = params;
  float* dP = smem;
  for (int i = threadIdx.x; i < pfloats; i += blockDim.x) dP[i] = 0.f;
  float* base = smem + pfloats;
  if (params_in_smem) {
    for (int i = threadIdx.x; i < S.n_params; i += blockDim.x) base[i] = params[i];
    P = base;
    base += pfloats;
  }
  __syncthreads();
  const int per_warp = (S.total + 2 * S.max_dim + 2 * S.c_in) * LP_LS;
  float* arena = base + warp * per_warp;
  float* gA = arena + S.total * LP_LS;
  float* gB = gA + S.max_dim * LP_LS;
  float* feat = gB + S.max_dim * LP_LS;
  float* gacc = feat + S.c_in * LP_LS;
  const int ray = (blockIdx.x * nwarps + warp) * LP_WARP + lane;
  const LpRayState s = lp_load_ray(R, ray, GG.g[0].B);
  const int rr = s.active ? ray : R.n - 1;
  const float vm = s.active ? (valid ? valid[rr] : 1.f) : 0.f;
  for (int k = 0; k < S.c_in; ++k) {
    feat[k * LP_LS + lane] = R.enc[(long long)rr * S.c_in + k];
    gacc[k * LP_LS + lane] = 0.f;
  }
  const int tot = M.S + M.S_inf;
  for (int step = 0; step < tot; ++step) {
    const float depth = lp_depth(step, s.near, s.far, M.S, M.S_inf, M.disparity_at_inf);
    float x = s.ox + depth * s.dx, y = s.oy + depth * s.dy, z = s.oz + depth * s.dz;
    if (M.contract) lp_contract(x, y, z);
    const float oob = M.mask_oob ? lp_in_bounds(x, y, z) : 1.f;
    lp_eval_splat_mlp(S, P, arena, feat, IN, s.b, x, y, z, oob, lane);
    // upstream gradient of the MLP output = sample(grad_grid) * valid  (splatter_bw.py:330-343)
    lp_lane_sample(GG, s.b, x, y, z, oob * vm, false, gA, lane);
    float* d_in = lp_mlp_backward(S.mlp, P, dP, arena, arena + S.xin * LP_LS, S.y, gA, gB, lane);
    for (int k = 0; k < S.c_in; ++k) gacc[k * LP_LS + lane] += d_in[k * LP_LS + lane];
    lp_lane_splat(IN, g_in, s.b, x, y, z, oob, d_in, lane);
  }
  if (s.active)
    for (int k = 0; k < S.c_in; ++k) g_feat[(long long)ray * S.c_in + k] = gacc[k * LP_LS + lane];
  __syncthreads();
  for (int i = threadIdx.x; i < S.n_params; i += blockDim.x) {
    const float v = dP[i];
    if (v != 0.f) lp_red_add1(g_params + i, v);
  }
}
