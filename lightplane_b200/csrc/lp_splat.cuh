// Splatter kernels.
//
// Plain splatter (no MLP): HBM/L2-atomic bound byte work.  A sub-warp of `lpr` lanes owns one ray
// and each lane a float4 chunk of the channels, so every tap of every sample is one coalesced
// `red.global.add.v4.f32` row segment (forward) or one 16-byte gather per lane (backward); the lanes
// of the sub-warp share the march (one lane per sample works out the taps, shuffles hand them round); feature
// and weight grid are accumulated in ONE march (the reference launches its kernel twice,
// lightplane_splatter.py:505,539).  Semantics: splatter_fw.py:71-165, splatter_bw.py:75-180.
//
// MLP splatter: generic lane-per-ray kernels built from the same blocks as the generic renderer
// (splatter_fw.py:168-309, splatter_bw.py:183-394).
#pragma once

#include "lp_render_generic.cuh"

// Taps of one grid as 32-bit ROW indices into the flat [rows, C] tensor (the host checks rows < 2^31) and weights;
// same corner order and zero-padding rule as lp_taps.  Returns the tap count (8 voxel / 4 plane) and whether any
// tap has a non-zero weight.
LP_DEVICE int lp_taps_rows(const LpGrid& g, int row_base, int b, float x, float y, float z, int* row, float* w,
                           bool& any, int& cell) {
  any = false;
  cell = -1;
  if (g.kind == LP_VOXEL) {
    float x0, fx, y0, fy, z0, fz;
    lp_axis(x, g.W, x0, fx);
    lp_axis(y, g.H, y0, fy);
    lp_axis(z, g.D, z0, fz);
    const int bb = row_base + b * g.D * g.H * g.W;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float wx, wy, wz;
      int ix, iy, iz;
      lp_corner(x0, fx, c & 1, g.W, wx, ix);
      lp_corner(y0, fy, (c >> 1) & 1, g.H, wy, iy);
      lp_corner(z0, fz, (c >> 2) & 1, g.D, wz, iz);
      w[c] = wx * wy * wz;
      any |= w[c] != 0.f;
      row[c] = bb + (iz * g.H + iy) * g.W + ix;
    }
    return 8;
  }
  float u, v;
  int U, V;
  if (g.kind == LP_PLANE_XY) { u = x; v = y; U = g.W; V = g.H; }
  else if (g.kind == LP_PLANE_XZ) { u = x; v = z; U = g.W; V = g.D; }
  else { u = y; v = z; U = g.H; V = g.D; }
  float u0, fu, v0, fv;
  lp_axis(u, U, u0, fu);
  lp_axis(v, V, v0, fv);
  const int bb = row_base + b * U * V;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float wu, wv;
    int iu, iv;
    lp_corner(u0, fu, c & 1, U, wu, iu);
    lp_corner(v0, fv, (c >> 1) & 1, V, wv, iv);
    w[c] = wu * wv;
    any |= w[c] != 0.f;
    row[c] = bb + iv * U + iu;
  }
#pragma unroll
  for (int c = 4; c < 8; ++c) { w[c] = 0.f; row[c] = 0; }
  if (any) cell = ((int)v0 + 1) * (U + 1) + (int)u0 + 1;  // the (unclamped) lower-corner texel: equal cell <=> equal rows
  return 4;
}

// The march of one ray is shared by the `lpr` lanes of its sub-warp: in every round lane j works out the position and
// the taps of step (round * lpr + j) -- the index arithmetic is done once per sample, not once per lane -- a ballot
// names the steps of the round that touch the grid at all, and for each of those the owning lane's row indices and
// weights are broadcast by shuffles while every lane reduces (forward) or gathers (backward) its own float4 channel
// chunks of the row.  Steps outside the grid (most of a scene view's samples) cost nothing beyond their share of the
// tap arithmetic.  On planes, consecutive samples of a ray that fall into the same texel cell (the usual case once the
// sample spacing is below the texel size) are merged first: the splatted feature is the ray's, so their four weights
// are summed over the run (a segmented suffix sum across the sub-warp) and the run is reduced once.  Forward only: with
// the gradient grid resident in L2 the backward's gathers are cheaper than the merge (measured, 128^2 x 32 triplane,
// 256 samples: forward 56 -> 44 ms, backward 26 -> 31 ms).
struct LpSplatLane {
  int sub, base;      // lane within the sub-warp, first lane of the sub-warp
  unsigned mask;      // the sub-warp's lanes
  long long ray;
  float ox, oy, oz, dx, dy, dz, near, far, vm;
  int b;
};
LP_DEVICE bool lp_splat_lane(const LpRays& R, const LpGridSet& G, const float* valid, int lpr, LpSplatLane& L) {
  const int lane = threadIdx.x & 31;
  L.sub = lane % lpr;
  L.base = lane - L.sub;
  L.mask = lpr >= 32 ? 0xffffffffu : (((1u << lpr) - 1u) << L.base);
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  L.ray = warp_global * (LP_WARP / lpr) + lane / lpr;
  if (L.ray >= R.n) return false;  // whole sub-warps leave together; no block-level barriers below
  const long long ray = L.ray;
  L.vm = valid ? valid[ray] : 1.f;
  L.ox = R.org[3 * ray]; L.oy = R.org[3 * ray + 1]; L.oz = R.org[3 * ray + 2];
  L.dx = R.dir[3 * ray]; L.dy = R.dir[3 * ray + 1]; L.dz = R.dir[3 * ray + 2];
  L.near = R.near[ray]; L.far = R.far[ray];
  L.b = min(max(R.gidx[ray], 0), G.g[0].B - 1);
  return true;
}
// position of this lane's step of the round; false when the step does not exist or is masked out
LP_DEVICE bool lp_splat_point(const LpSplatLane& L, const LpMarch& M, int step, float& x, float& y, float& z) {
  const int tot = M.S + M.S_inf;
  const float depth = lp_depth(min(step, tot - 1), L.near, L.far, M.S, M.S_inf, M.disparity_at_inf);
  x = L.ox + depth * L.dx; y = L.oy + depth * L.dy; z = L.oz + depth * L.dz;
  if (M.contract) lp_contract(x, y, z);
  return step < tot && !(M.mask_oob && lp_in_bounds(x, y, z) == 0.f);
}

// Ballot of the steps of this round that have work, after folding runs of equal `cell` into their first lane.
template <bool MERGE>
LP_DEVICE unsigned lp_splat_todo(const LpSplatLane& L, int lpr, bool live, bool any, int cell, int nt, float* w) {
  const bool work = live && any;
  if (!MERGE || nt == 8 || lpr == 1) return (__ballot_sync(L.mask, work) & L.mask) >> L.base;
  const int key = work ? cell : (int)(0x80000000u | (unsigned)L.sub);  // lanes without work never merge
  const int prev = __shfl_sync(L.mask, key, L.sub - 1, lpr);
  const bool head = L.sub == 0 || prev != key;
  const unsigned heads = (__ballot_sync(L.mask, head) & L.mask) >> L.base;
  const unsigned above = L.sub == 31 ? 0u : heads >> (L.sub + 1);
  const int end = above ? L.sub + __ffs((int)above) - 1 : lpr - 1;  // last lane of this lane's run
  for (int d = 1; d < lpr; d <<= 1) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float o = __shfl_sync(L.mask, w[t], L.sub + d, lpr);
      if (L.sub + d <= end) w[t] += o;
    }
  }
  return (__ballot_sync(L.mask, head && work) & L.mask) >> L.base;
}

template <int VPL>
__global__ void lp_splat_fwd_kernel(LpRays R, LpMarch M, LpGridSet OUT, float* __restrict__ weight,
                                    const float* __restrict__ valid, int lpr) {
  LpSplatLane L;
  if (!lp_splat_lane(R, OUT, valid, lpr, L)) return;
  const int C = OUT.C;
  float4 f[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    f[v] = lp_ldg4(R.enc + L.ray * C + 4 * (L.sub + lpr * v));
    f[v].x *= L.vm; f[v].y *= L.vm; f[v].z *= L.vm; f[v].w *= L.vm;
  }
  const bool splat_weight = L.sub == 0 && weight != nullptr && L.vm != 0.f;
  const int tot = M.S + M.S_inf;
  for (int step0 = 0; step0 < tot; step0 += lpr) {
    float x, y, z;
    const bool live = lp_splat_point(L, M, step0 + L.sub, x, y, z);
    for (int gi = 0; gi < OUT.n; ++gi) {
      int row[8];
      float w[8];
      bool any;
      int cell;
      const int nt = lp_taps_rows(OUT.g[gi], (int)(OUT.g[gi].base / C), L.b, x, y, z, row, w, any, cell);
      unsigned todo = lp_splat_todo<true>(L, lpr, live, any, cell, nt, w);
      while (todo) {
        const int j = __ffs((int)todo) - 1;
        todo &= todo - 1;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (t < nt) {
            const float wt = __shfl_sync(L.mask, w[t], j, lpr);
            const int rt = __shfl_sync(L.mask, row[t], j, lpr);
            if (wt != 0.f) {
              float* dst = OUT.data + (long long)rt * C + 4 * L.sub;
#pragma unroll
              for (int v = 0; v < VPL; ++v)
                lp_red_add4(dst + 4 * lpr * v, wt * f[v].x, wt * f[v].y, wt * f[v].z, wt * f[v].w);
              if (splat_weight) lp_red_add1(weight + rt, wt * L.vm);
            }
          }
        }
      }
    }
  }
}

template <int VPL>
__global__ void lp_splat_bwd_kernel(LpRays R, LpMarch M, LpGridSet GG, const float* __restrict__ valid,
                                    float* __restrict__ g_feat, int lpr) {
  LpSplatLane L;
  if (!lp_splat_lane(R, GG, valid, lpr, L)) return;
  const int C = GG.C;
  float4 acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int tot = M.S + M.S_inf;
  for (int step0 = 0; step0 < tot; step0 += lpr) {
    float x, y, z;
    const bool live = lp_splat_point(L, M, step0 + L.sub, x, y, z);
    for (int gi = 0; gi < GG.n; ++gi) {
      int row[8];
      float w[8];
      bool any;
      int cell;
      const int nt = lp_taps_rows(GG.g[gi], (int)(GG.g[gi].base / C), L.b, x, y, z, row, w, any, cell);
      unsigned todo = lp_splat_todo<false>(L, lpr, live, any, cell, nt, w);
      while (todo) {
        const int j = __ffs((int)todo) - 1;
        todo &= todo - 1;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (t < nt) {
            const float wt = __shfl_sync(L.mask, w[t], j, lpr);
            const int rt = __shfl_sync(L.mask, row[t], j, lpr);
            if (wt != 0.f) {
              const float* src = GG.data + (long long)rt * C + 4 * L.sub;
#pragma unroll
              for (int v = 0; v < VPL; ++v) {
                const float4 g = lp_ldg4(src + 4 * lpr * v);
                acc[v].x = fmaf(wt, g.x, acc[v].x); acc[v].y = fmaf(wt, g.y, acc[v].y);
                acc[v].z = fmaf(wt, g.z, acc[v].z); acc[v].w = fmaf(wt, g.w, acc[v].w);
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    float* o = g_feat + L.ray * C + 4 * (L.sub + lpr * v);
    o[0] = acc[v].x * L.vm; o[1] = acc[v].y * L.vm; o[2] = acc[v].z * L.vm; o[3] = acc[v].w * L.vm;
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
