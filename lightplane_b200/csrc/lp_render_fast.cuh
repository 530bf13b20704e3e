// Specialised Renderer kernels for the default decoder shape (trunk/opacity/colour = 2/2/2 layers,
// hidden width 32, C in {16,32} grid channels, <= 3 colour channels, no colour grid / scaffold):
// the per-sample MLP runs on the tensor cores.
//
// Mapping.  One warp marches 32 rays in lock-step (two m16 tiles of `mma.sync.m16n8k8`, TF32
// operands, FP32 accumulate).  Lane (g = lane>>2, t = lane&3) owns rows g and g+8 of each tile, so
// a quad of 4 lanes shares 4 rays and each lane gathers a float4 channel chunk of every tap.
// Activations never leave registers between layers: the C-fragment of layer l (row g, cols 2t,2t+1)
// IS the A-fragment of layer l+1 once the K index is permuted (k-slot t <-> col 2t, slot t+4 <-> col
// 2t+1); the permutation is folded into the shared-memory image of the weights, which is stored in
// fragment order so that every B-fragment is one conflict-free LDS.128.
//
// Precision.  The reference computes in IEEE fp32 (triton_src/shared/const.py:8-9).  Plain TF32
// misses the 1e-3 gradient bar by ~10x, so forward and recompute use the 3xTF32
// split  x*w ~= x_lo*w_hi + x_hi*w_lo + x_hi*w_hi  (hi = cvt.rna.tf32, lo = exact remainder);
// the backward products (dX, dW) use single TF32 with round-to-nearest operands.
//
// Semantics are those of lp_render_generic.cuh (which restates renderer_fw.py / renderer_bw.py).
#pragma once

#include "lp_render_generic.cuh"

namespace lpf {

constexpr int H = 32;
#ifndef LP_GATHER_STREAM
#define LP_GATHER_STREAM 0  // backward gathers bypass L1 allocation (keeps the ray encodings L1-resident)
#endif
#ifndef LP_GATHER_SKIP
#define LP_GATHER_SKIP 1  // skip the loads of a grid the sample misses entirely (mostly warp-uniform for camera rays)
#endif
#ifndef LP_ALIGN_MASK
#define LP_ALIGN_MASK 0  // re-align the warps of a CTA every (mask+1) steps (I-cache sharing vs barrier stalls)
#endif

template <int C>
struct Lay {
  static constexpr int KS0 = C / 8;  // k-steps of the first trunk layer
  // ---- forward image: fragment-ordered {hi0, hi1, lo0, lo1} per (k-step, n-tile, lane) ----
  static constexpr int F_T0 = 0;
  static constexpr int F_T1 = F_T0 + KS0 * 4 * 128;
  static constexpr int F_O0 = F_T1 + 2048;
  static constexpr int F_C0 = F_O0 + 2048;
  static constexpr int F_LC = F_C0 + 2048;  // last layer, colour columns (one n-tile)
  static constexpr int F_LO = F_LC + 512;   // last layer, opacity columns
  static constexpr int BIAS = F_LO + 512;   // b_t0[32] b_t1[32] b_o0[32] b_c0[32] b_last[8]
  static constexpr int FWD_END = BIAS + 136;
  // ---- dX image: fragment-ordered {hi0, hi1} per (k-step, n-tile, lane) ----
  static constexpr int X_LC = FWD_END;      // d_hc = dY_last * Wc1^T     [1][4]
  static constexpr int X_LO = X_LC + 256;   // d_ho = dY_last * wo1^T     [1][4]
  static constexpr int X_C0 = X_LO + 256;   // [4][4]
  static constexpr int X_O0 = X_C0 + 1024;
  static constexpr int X_T1 = X_O0 + 1024;
  static constexpr int X_T0 = X_T1 + 1024;  // [4][C/8]
  static constexpr int END = X_T0 + 4 * (C / 8) * 64;
};

// input row of W (in-feature) feeding k-slot `slot` (0..7) of k-step j
template <int C>
LP_DEVICE int lp_t0_row(int j, int slot) {
  const int tt = slot & 3, e = slot >> 2;
  if (C == 16) return 4 * tt + 2 * j + e;
  return 16 * (j >> 1) + 4 * tt + 2 * (j & 1) + e;
}
LP_DEVICE int lp_std_row(int j, int slot) { return 8 * j + 2 * (slot & 3) + (slot >> 2); }
// grid channel held in column `col` (0..7) of n-tile nn of d_x0
template <int C>
LP_DEVICE int lp_t0_chan(int nn, int col) {
  const int tt = col >> 1, e = col & 1;
  if (C == 16) return 4 * tt + 2 * nn + e;
  return 16 * (nn >> 1) + 4 * tt + 2 * (nn & 1) + e;
}

// Build the shared-memory weight image from the flat parameter vector (all threads of the CTA).
template <int C, bool WITH_DX>
LP_DEVICE void lp_build_weights(float* sm, const float* __restrict__ P, const LpDecoder& D) {
  using L = Lay<C>;
  const LpLayer &t0 = D.trunk.l[0], &t1 = D.trunk.l[1], &o0 = D.opacity.l[0], &o1 = D.opacity.l[1],
                &c0 = D.color.l[0], &c1 = D.color.l[1];
  const int tid = threadIdx.x, nth = blockDim.x;
  auto put4 = [&](float* dst, float wa, float wb) {
    const float ha = lp_tf32_rna(wa), hb = lp_tf32_rna(wb);
    dst[0] = ha; dst[1] = hb; dst[2] = lp_tf32_rna(wa - ha); dst[3] = lp_tf32_rna(wb - hb);
  };
  // hidden-size layers with C-fragment inputs
  for (int which = 0; which < 3; ++which) {
    const LpLayer& Ly = which == 0 ? t1 : (which == 1 ? o0 : c0);
    float* base = sm + (which == 0 ? L::F_T1 : (which == 1 ? L::F_O0 : L::F_C0));
    for (int e = tid; e < 4 * 4 * 32; e += nth) {
      const int lane = e & 31, n = (e >> 5) & 3, j = e >> 7, g = lane >> 2, t = lane & 3;
      const float* W = P + Ly.w_off;
      put4(base + e * 4, W[lp_std_row(j, t) * Ly.N + 8 * n + g], W[lp_std_row(j, t + 4) * Ly.N + 8 * n + g]);
    }
  }
  for (int e = tid; e < L::KS0 * 4 * 32; e += nth) {
    const int lane = e & 31, n = (e >> 5) & 3, j = e >> 7, g = lane >> 2, t = lane & 3;
    const float* W = P + t0.w_off;
    put4(sm + L::F_T0 + e * 4, W[lp_t0_row<C>(j, t) * t0.N + 8 * n + g], W[lp_t0_row<C>(j, t + 4) * t0.N + 8 * n + g]);
  }
  // last layer: one n-tile, columns [c0, o, c1, o, c2, o, -, o]
  for (int e = tid; e < 4 * 32; e += nth) {
    const int lane = e & 31, j = e >> 5, g = lane >> 2, t = lane & 3;
    const int ra = lp_std_row(j, t), rb = lp_std_row(j, t + 4);
    const bool is_col = ((g & 1) == 0) && ((g >> 1) < D.n_feat);
    const float* Wc = P + c1.w_off;
    const float* Wo = P + o1.w_off;
    put4(sm + L::F_LC + e * 4, is_col ? Wc[ra * c1.N + (g >> 1)] : 0.f, is_col ? Wc[rb * c1.N + (g >> 1)] : 0.f);
    put4(sm + L::F_LO + e * 4, (g & 1) ? Wo[ra] : 0.f, (g & 1) ? Wo[rb] : 0.f);
  }
  for (int e = tid; e < 136; e += nth) {
    float v;
    if (e < 32) v = P[t0.b_off + e];
    else if (e < 64) v = P[t1.b_off + e - 32];
    else if (e < 96) v = P[o0.b_off + e - 64];
    else if (e < 128) v = P[c0.b_off + e - 96];
    else {
      const int col = e - 128;
      v = (col & 1) ? P[o1.b_off] : ((col >> 1) < D.n_feat ? P[c1.b_off + (col >> 1)] : 0.f);
    }
    sm[L::BIAS + e] = v;
  }
  if (WITH_DX) {
    // dX B-fragments: b0 = W[in(n-tile nn, g)][8jk+2t], b1 = W[...][8jk+2t+1]
    for (int which = 0; which < 3; ++which) {
      const LpLayer& Ly = which == 0 ? c0 : (which == 1 ? o0 : t1);
      float* base = sm + (which == 0 ? L::X_C0 : (which == 1 ? L::X_O0 : L::X_T1));
      for (int e = tid; e < 4 * 4 * 32; e += nth) {
        const int lane = e & 31, nn = (e >> 5) & 3, jk = e >> 7, g = lane >> 2, t = lane & 3;
        const float* W = P + Ly.w_off + (8 * nn + g) * Ly.N + 8 * jk + 2 * t;
        base[e * 2] = lp_tf32_rna(W[0]);
        base[e * 2 + 1] = lp_tf32_rna(W[1]);
      }
    }
    for (int e = tid; e < 4 * (C / 8) * 32; e += nth) {
      const int lane = e & 31, nn = (e >> 5) % (C / 8), jk = (e >> 5) / (C / 8), g = lane >> 2, t = lane & 3;
      const float* W = P + t0.w_off + lp_t0_chan<C>(nn, g) * t0.N + 8 * jk + 2 * t;
      sm[L::X_T0 + e * 2] = lp_tf32_rna(W[0]);
      sm[L::X_T0 + e * 2 + 1] = lp_tf32_rna(W[1]);
    }
    for (int e = tid; e < 4 * 32; e += nth) {  // last layer: k-slot t <-> colour t, slot 4 <-> opacity
      const int lane = e & 31, nn = e >> 5, g = lane >> 2, t = lane & 3;
      const int i = 8 * nn + g;
      sm[L::X_LC + e * 2] = (t < D.n_feat) ? lp_tf32_rna(P[c1.w_off + i * c1.N + t]) : 0.f;
      sm[L::X_LC + e * 2 + 1] = 0.f;
      sm[L::X_LO + e * 2] = 0.f;
      sm[L::X_LO + e * 2 + 1] = (t == 0) ? lp_tf32_rna(P[o1.w_off + i]) : 0.f;
    }
  }
}

// acc[mt][n][.] <- bias of columns 8n+2t, 8n+2t+1
LP_DEVICE void lp_init_bias(float (&acc)[2][4][4], const float* bias, int t) {
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const float b0 = bias[8 * n + 2 * t], b1 = bias[8 * n + 2 * t + 1];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) { acc[mt][n][0] = b0; acc[mt][n][1] = b1; acc[mt][n][2] = b0; acc[mt][n][3] = b1; }
  }
}

// One 3xTF32 dense layer for both m-tiles: acc[mt][n] += A[mt] * W, A given in A-fragment order.
// The three products of one accumulator are dependent; they are issued product-major so that
// eight independent accumulators separate two dependent mma instructions.
template <int KSTEPS>
LP_DEVICE void lp_layer3x(const float* Wf, float (&acc)[2][4][4], const float (&ain)[2][KSTEPS][4], int lane) {
#pragma unroll
  for (int j = 0; j < KSTEPS; ++j) {
    float ahi[2][4], alo[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) { ahi[mt][i] = lp_tf32_rna(ain[mt][j][i]); alo[mt][i] = ain[mt][j][i] - ahi[mt][i]; }
    float bh[4][2], bl[4][2];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float4 w = *reinterpret_cast<const float4*>(Wf + ((j * 4 + n) * 32 + lane) * 4);
      bh[n][0] = w.x; bh[n][1] = w.y; bl[n][0] = w.z; bl[n][1] = w.w;
    }
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) lp_mma_tf32(acc[mt][n], alo[mt], bh[n]);
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) lp_mma_tf32(acc[mt][n], ahi[mt], bl[n]);
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) lp_mma_tf32(acc[mt][n], ahi[mt], bh[n]);
  }
}

// single n-tile variant (last layer)
LP_DEVICE void lp_layer3x_n1(const float* Wf, float (&acc)[2][4], const float (&ain)[2][4][4], int lane) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 w = *reinterpret_cast<const float4*>(Wf + (j * 32 + lane) * 4);
    const float bh[2] = {w.x, w.y}, bl[2] = {w.z, w.w};
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float ahi[4], alo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { ahi[i] = lp_tf32_rna(ain[mt][j][i]); alo[i] = ain[mt][j][i] - ahi[i]; }
      lp_mma_tf32(acc[mt], alo, bh);
      lp_mma_tf32(acc[mt], ahi, bl);
      lp_mma_tf32(acc[mt], ahi, bh);
    }
  }
}

// ReLU the C-fragments of a layer and re-label them as the next layer's A-fragments
// (a0=(g,slot t)=c0, a1=(g+8,slot t)=c2, a2=(g,slot t+4)=c1, a3=(g+8,slot t+4)=c3).
LP_DEVICE void lp_relu_to_a(const float (&acc)[2][4][4], float (&a)[2][4][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      a[mt][n][0] = fmaxf(acc[mt][n][0], 0.f);
      a[mt][n][1] = fmaxf(acc[mt][n][2], 0.f);
      a[mt][n][2] = fmaxf(acc[mt][n][1], 0.f);
      a[mt][n][3] = fmaxf(acc[mt][n][3], 0.f);
    }
}

// -------------------------------------------------------------------------------------------
// Ray geometry.  Sampling positions and taps are computed by the lane that OWNS the ray (lane r <->
// ray rbase + r); the MLP works on quad-shared rows (row i of lane (g,t) = ray 16*(i>>1)+8*(i&1)+g).
// Sampled features / their gradients cross between the two layouts through a small swizzled
// shared-memory tile, so the tap arithmetic is done once per ray instead of once per lane of a quad.
// -------------------------------------------------------------------------------------------
struct Ray1 {
  float ox, oy, oz, dx, dy, dz, near, far;
  int b, ray;
  bool active;
};
LP_DEVICE Ray1 lp_load_ray1(const LpRays& R, int ray, int batch) {
  Ray1 r;
  r.active = ray < R.n;
  r.ray = ray;
  const int q = r.active ? ray : R.n - 1;
  r.ox = R.org[3 * q]; r.oy = R.org[3 * q + 1]; r.oz = R.org[3 * q + 2];
  r.dx = R.dir[3 * q]; r.dy = R.dir[3 * q + 1]; r.dz = R.dir[3 * q + 2];
  r.near = R.near[q]; r.far = R.far[q];
  r.b = min(max(R.gidx[q], 0), batch - 1);
  return r;
}

// warp-uniform depth schedule of one step: depth = a + b*c  with per-ray (a, b) chosen by `inf`
struct Sched {
  float cur, prev;  // regular: j/(S-1), (j-1)/(S-1);  background: 1/n_disp(k), 1/n_disp(k-1)
  bool inf, first_inf, single;
};
LP_DEVICE Sched lp_sched(int step, const LpMarch& M) {
  Sched s;
  s.inf = step >= M.S;
  s.single = M.S <= 1;
  s.first_inf = step == M.S;
  if (!s.inf) {
    const float inv = s.single ? 0.f : 1.f / (float)(M.S - 1);
    s.cur = (float)step * inv;
    s.prev = (float)(step - 1) * inv;
  } else {
    const int k = step - M.S;
    auto sc = [&](int kk) {  // 1 / ((1-f) + d_inf*f), f = (kk+1)/S_inf  (see lp_depth)
      const float f = (float)(kk + 1) / (float)M.S_inf;
      const float omf = (float)(M.S_inf - (kk + 1)) / (float)M.S_inf;
      return 1.f / (omf + M.disparity_at_inf * f);
    };
    s.cur = sc(k);
    s.prev = sc(k - 1);
  }
  return s;
}
LP_DEVICE void lp_depth_delta(const Sched& s, float near, float far, float& depth, float& delta) {
  if (!s.inf) {
    if (s.single) { depth = near; delta = 1.f; return; }
    depth = (far - near) * s.cur + near;
    delta = depth - ((far - near) * s.prev + near);
  } else {
    depth = far * s.cur;
    delta = depth - (s.first_inf ? ((far - near) * 1.f + near) : far * s.prev);
  }
}

// taps of one grid with 32-bit element offsets (the fast path requires < 2^31 grid elements)
LP_DEVICE void lp_axis_i(float p, int size, int& i0, float& frac) {
  float i = ((p + 1.f) * 0.5f) * (float)size - 0.5f;
  if (size <= 1) i = 0.f;
  const float f0 = floorf(i);
  frac = i - f0;
  i0 = (int)fminf(fmaxf(f0, -2.f), (float)size);  // clamp keeps the int conversion defined
}
LP_DEVICE void lp_corner_i(int i0, float frac, int size, float& w0, float& w1, int& c0, int& c1) {
  w0 = ((unsigned)i0 < (unsigned)size) ? 1.f - frac : 0.f;
  w1 = ((unsigned)(i0 + 1) < (unsigned)size) ? frac : 0.f;
  c0 = min(max(i0, 0), size - 1);
  c1 = min(max(i0 + 1, 0), size - 1);
}
LP_DEVICE int lp_taps_i32(const LpGrid& g, int C, int b, float x, float y, float z, int* off, float* w) {
  if (g.kind == LP_VOXEL) {
    int x0, y0, z0, cx[2], cy[2], cz[2];
    float fx, fy, fz, wx[2], wy[2], wz[2];
    lp_axis_i(x, g.W, x0, fx); lp_axis_i(y, g.H, y0, fy); lp_axis_i(z, g.D, z0, fz);
    lp_corner_i(x0, fx, g.W, wx[0], wx[1], cx[0], cx[1]);
    lp_corner_i(y0, fy, g.H, wy[0], wy[1], cy[0], cy[1]);
    lp_corner_i(z0, fz, g.D, wz[0], wz[1], cz[0], cz[1]);
    const int bbase = (int)g.base + b * g.D * g.H * g.W * C;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      w[c] = wx[c & 1] * wy[(c >> 1) & 1] * wz[(c >> 2) & 1];
      off[c] = bbase + ((cz[(c >> 2) & 1] * g.H + cy[(c >> 1) & 1]) * g.W + cx[c & 1]) * C;
    }
    return 8;
  }
  float u, v;
  int U, V;
  if (g.kind == LP_PLANE_XY) { u = x; v = y; U = g.W; V = g.H; }
  else if (g.kind == LP_PLANE_XZ) { u = x; v = z; U = g.W; V = g.D; }
  else { u = y; v = z; U = g.H; V = g.D; }
  int u0, v0, cu[2], cv[2];
  float fu, fv, wu[2], wv[2];
  lp_axis_i(u, U, u0, fu); lp_axis_i(v, V, v0, fv);
  lp_corner_i(u0, fu, U, wu[0], wu[1], cu[0], cu[1]);
  lp_corner_i(v0, fv, V, wv[0], wv[1], cv[0], cv[1]);
  const int bbase = (int)g.base + b * U * V * C;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    w[c] = wu[c & 1] * wv[c >> 1];
    off[c] = bbase + (cv[c >> 1] * U + cu[c & 1]) * C;
  }
  return 4;
}

// float index of (ray row r, float4 chunk k) inside the per-warp [32][C] transfer tile; the XOR
// makes both the row-wise float4 accesses of the ray-owner lanes and the chunk-wise accesses of
// the quad lanes bank-conflict free
template <int C>
LP_DEVICE int lp_xs(int r, int k) {
  if (C == 16) return r * 16 + ((k ^ ((r >> 1) & 3)) << 2);
  return r * 32 + ((k ^ (((r & 1) << 2) | ((r >> 1) & 3))) << 2);
}

// The ray-owner lane samples all C channels of its sample point and drops them into the tile.
template <int C, bool STREAM = false>
LP_DEVICE void lp_gather_lane(const LpGridSet& G, int b, float x, float y, float z, float oob, float* xs, int lane) {
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  for (int gi = 0; gi < G.n; ++gi) {
    int off[8];
    float w[8];
    const int nt = lp_taps_i32(G.g[gi], C, b, x, y, z, off, w);
#if LP_GATHER_SKIP
    float wsum = 0.f;
#pragma unroll
    for (int tp = 0; tp < 8; ++tp)
      if (tp < nt) wsum += w[tp];
    if (wsum == 0.f) continue;  // the sample misses this grid entirely
#endif
#pragma unroll
    for (int tp = 0; tp < 8; ++tp) {
      if (tp < nt) {  // zero-weight taps carry clamped (valid) addresses: load unconditionally, no branches
#pragma unroll
        for (int k = 0; k < C / 4; ++k) {
          const float4 v = (STREAM && LP_GATHER_STREAM) ? lp_ldg4_stream(G.data + off[tp] + 4 * k) : lp_ldg4(G.data + off[tp] + 4 * k);
          acc[4 * k] = fmaf(w[tp], v.x, acc[4 * k]); acc[4 * k + 1] = fmaf(w[tp], v.y, acc[4 * k + 1]);
          acc[4 * k + 2] = fmaf(w[tp], v.z, acc[4 * k + 2]); acc[4 * k + 3] = fmaf(w[tp], v.w, acc[4 * k + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < C / 4; ++k)
    *reinterpret_cast<float4*>(xs + lp_xs<C>(lane, k)) =
        make_float4(acc[4 * k] * oob, acc[4 * k + 1] * oob, acc[4 * k + 2] * oob, acc[4 * k + 3] * oob);
}
// quad lane (g,t): channel chunk(s) t (and 4+t) of its four rows
template <int C>
LP_DEVICE void lp_read_rows(const float* xs, int g, int t, float (&xa)[4][C / 4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < C / 16; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(xs + lp_xs<C>(8 * i + g, 4 * k + t));
      xa[i][4 * k] = v.x; xa[i][4 * k + 1] = v.y; xa[i][4 * k + 2] = v.z; xa[i][4 * k + 3] = v.w;
    }
}
// adjoint: the ray-owner lane scatters its row of the tile into the grid gradient.  Most samples of a
// camera ray miss a given plane (or the volume) entirely, so the products and the reductions of a
// grid are skipped when none of its taps carries weight.
// (A warp-level pre-aggregation of the 32 rays' overlapping footprints on the tensor core -- one
// reduction per touched texel instead of one per ray and tap -- was built and measured: 8x fewer L2
// reductions but ~450 extra instructions per step, and the kernel is issue-bound, not L2-atomic
// bound: 169.6 ms vs 164.5 ms per step.  See DESIGN.md section 5.)
template <int C>
LP_DEVICE void lp_splat_lane(const LpGridSet& G, float* grad, int b, float x, float y, float z, float oob,
                             const float* xs, int lane) {
  if (oob == 0.f) return;
  float d[C];
#pragma unroll
  for (int k = 0; k < C / 4; ++k) {
    const float4 v = *reinterpret_cast<const float4*>(xs + lp_xs<C>(lane, k));
    d[4 * k] = v.x * oob; d[4 * k + 1] = v.y * oob; d[4 * k + 2] = v.z * oob; d[4 * k + 3] = v.w * oob;
  }
  for (int gi = 0; gi < G.n; ++gi) {
    int off[8];
    float w[8];
    const int nt = lp_taps_i32(G.g[gi], C, b, x, y, z, off, w);
    float wsum = 0.f;  // weights are >= 0
#pragma unroll
    for (int tp = 0; tp < 8; ++tp)
      if (tp < nt) wsum += w[tp];
    if (wsum == 0.f) continue;
#pragma unroll
    for (int tp = 0; tp < 8; ++tp) {
      if (tp < nt) {
        const bool on = w[tp] != 0.f;  // taps outside the grid: no atomic traffic
#pragma unroll
        for (int k = 0; k < C / 4; ++k)
          lp_red_add4_if(on, grad + off[tp] + 4 * k, w[tp] * d[4 * k], w[tp] * d[4 * k + 1], w[tp] * d[4 * k + 2],
                         w[tp] * d[4 * k + 3]);
      }
    }
  }
}

// x0 rows -> A-fragments of the first trunk layer
template <int C>
LP_DEVICE void lp_x0_to_a(const float (&xa)[4][C / 4], float (&a)[2][C / 8][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int j = 0; j < C / 8; ++j) {
      const int f = (C == 16) ? 2 * j : 4 * (j >> 1) + 2 * (j & 1);
      a[mt][j][0] = xa[2 * mt][f];
      a[mt][j][1] = xa[2 * mt + 1][f];
      a[mt][j][2] = xa[2 * mt][f + 1];
      a[mt][j][3] = xa[2 * mt + 1][f + 1];
    }
}

// ===========================================================================================
// forward
// ===========================================================================================
template <int C>
__global__ void __launch_bounds__(256, 2) lp_render_fwd_fast_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G,
                                                                const float* __restrict__ params,
                                                                float* __restrict__ out_len,
                                                                float* __restrict__ out_nlt,
                                                                float* __restrict__ out_feat, int feat_stride) {
  using L = Lay<C>;
  LP_DYN_SMEM(float, smem);
  lp_build_weights<C, false>(smem, params, D);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* enc_s = smem + L::FWD_END + warp * (32 * 40 + 32 * C);  // per-warp [32 rays][40] encoding tile
  float* xs = enc_s + 32 * 40;                                    // per-warp [32][C] transfer tile
  const float* bias = smem + L::BIAS;
  const int num_tiles = (R.n + 31) / 32;
  const int tot = M.S + M.S_inf;

  // Every warp of every CTA runs the same number of tiles (surplus tiles are all-inactive rays), so
  // the per-step CTA barrier below is well formed; it keeps the warps on the same instruction-cache
  // lines -- the unrolled step body is ~100 KB of code and otherwise every warp streams it separately.
  const int tiles_per_warp = (num_tiles + gridDim.x * nwarps - 1) / (gridDim.x * nwarps);
  for (int it = 0; it < tiles_per_warp; ++it) {
    const int tile = (it * gridDim.x + blockIdx.x) * nwarps + warp;
    const int rbase = tile * 32;
    const Ray1 me = lp_load_ray1(R, rbase + lane, G.g[0].B);
    float rnear[4], rfar[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rnear[i] = __shfl_sync(LP_FULL_MASK, me.near, 8 * i + g);
      rfar[i] = __shfl_sync(LP_FULL_MASK, me.far, 8 * i + g);
    }
    __syncwarp();
    for (int e = lane; e < 32 * 8; e += 32) {  // 32 rays x 8 float4
      const int row = e >> 3, c4 = e & 7;
      const int q = min(rbase + row, R.n - 1);
      *reinterpret_cast<float4*>(enc_s + row * 40 + 4 * c4) = lp_ldg4(R.enc + (long long)q * H + 4 * c4);
    }
    __syncwarp();
    float nlt[4] = {0.f, 0.f, 0.f, 0.f}, T[4] = {1.f, 1.f, 1.f, 1.f}, accum[4] = {0.f, 0.f, 0.f, 0.f};

    for (int step = 0; step < tot; ++step) {
      if ((step & LP_ALIGN_MASK) == 0) __syncthreads();
      const Sched sc = lp_sched(step, M);
      float xa[4][C / 4], depth[4], delta[4];
      {
        float dep, del;
        lp_depth_delta(sc, me.near, me.far, dep, del);
        float x = me.ox + dep * me.dx, y = me.oy + dep * me.dy, z = me.oz + dep * me.dz;
        if (M.contract) lp_contract(x, y, z);
        const float oob = M.mask_oob ? lp_in_bounds(x, y, z) : 1.f;
        __syncwarp();  // previous step's readers are done with the tile
        lp_gather_lane<C>(G, me.b, x, y, z, oob, xs, lane);
        __syncwarp();
        lp_read_rows<C>(xs, g, t, xa);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) lp_depth_delta(sc, rnear[i], rfar[i], depth[i], delta[i]);
      float acc[2][4][4], a[2][4][4];
      {
        float a0[2][C / 8][4];
        lp_x0_to_a<C>(xa, a0);
        lp_init_bias(acc, bias, t);
        lp_layer3x<C / 8>(smem + L::F_T0, acc, a0, lane);
      }
      lp_relu_to_a(acc, a);
      lp_init_bias(acc, bias + 32, t);
      lp_layer3x<4>(smem + L::F_T1, acc, a, lane);
      float tr[2][4][4];  // trunk output (post-ReLU) in A-fragment order
      lp_relu_to_a(acc, tr);
      // opacity hidden
      lp_init_bias(acc, bias + 64, t);
      lp_layer3x<4>(smem + L::F_O0, acc, tr, lane);
      float last[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) { last[mt][0] = last[mt][2] = bias[128 + 2 * t]; last[mt][1] = last[mt][3] = bias[128 + 2 * t + 1]; }
      lp_relu_to_a(acc, a);
      lp_layer3x_n1(smem + L::F_LO, last, a, lane);
      // colour hidden: input = trunk + ray encoding
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const float2 e0 = *reinterpret_cast<const float2*>(enc_s + (16 * mt + g) * 40 + 8 * n + 2 * t);
          const float2 e1 = *reinterpret_cast<const float2*>(enc_s + (16 * mt + 8 + g) * 40 + 8 * n + 2 * t);
          tr[mt][n][0] += e0.x; tr[mt][n][2] += e0.y; tr[mt][n][1] += e1.x; tr[mt][n][3] += e1.y;
        }
      lp_init_bias(acc, bias + 96, t);
      lp_layer3x<4>(smem + L::F_C0, acc, tr, lane);
      lp_relu_to_a(acc, a);
      lp_layer3x_n1(smem + L::F_LC, last, a, lane);
      // ---- compositing: every lane tracks T of its 4 rows; lane t<3 owns colour t, lane 3 the length
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int mt = i >> 1, h = i & 1;
        float raw = last[mt][2 * h + 1];
        if (M.noise) raw += M.sigma * lp_sample_noise(M, rbase + 8 * i + g, step);
        nlt[i] += delta[i] * M.gain * lp_softplus(raw);
        const float Tn = expf(-nlt[i]);
        const float w = T[i] - Tn;
        T[i] = Tn;
        accum[i] = fmaf(w, (t == 3) ? depth[i] : lp_sigmoid(last[mt][2 * h]), accum[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ray = rbase + 8 * i + g;
      if (ray < R.n) {
        if (t < D.n_feat) out_feat[(long long)ray * feat_stride + t] = accum[i];
        if (t == 3) { out_len[ray] = accum[i]; out_nlt[ray] = nlt[i]; }
      }
    }
  }
}


// ===========================================================================================
// backward
// ===========================================================================================
// The register-resident chain (3xTF32 recompute, TF32 dX) stays on mma.sync; the parameter-gradient
// GEMMs  dW_l += X_l^T dY_l  -- a reduction over EVERY sample the CTA ever touches -- run on the
// 5th-generation tensor cores: each warp drops bf16 copies of its activations X_l and gradients dY_l
// (32 samples) into shared memory as MN-major UMMA operands and issues tcgen05.mma instructions that
// accumulate into ONE set of fp32 accumulators per CTA held in TMEM for the whole kernel (144 KB of
// accumulator state per CTA would otherwise live in registers / shared memory and cap the kernel at
// 4 warps per SM).  bf16 rounding of X / dY is unbiased and independent per sample, so it averages
// out over the millions of samples reduced into each dW entry.
//
// Operand stacks of one warp (rows = features, 32 samples each, see lp_platform.cuh for the layout):
//   A1 = [h1 (32) | trunk (32) | trunk+enc (32) | x0 (C, zero padded to 32)]        128 rows
//   A2 = [colour hidden (32) | opacity hidden (32) | ones (1) | ...]                128 rows (65 used)
//   B_j = dY of trunk layer 1, opacity hidden, colour hidden, trunk layer 0 (32 wide), B_last (16 wide)
// TMEM columns:  [0,128): A1 x B_j -> dW of the four 32-wide layers on the block diagonal;
//                [128,256): A2 x B_j -> row 64 (ones) = bias gradients;  [256,272): A2 x B_last.

LP_DEVICE int lp_sw32(int s, int f) { return s * 32 + (f ^ ((s & 3) << 3)); }

// ReLU mask of an activation in A-fragment order: bit (mt*16 + n*4 + i) set iff a[mt][n][i] > 0
LP_DEVICE unsigned lp_mask_a(const float (&a)[2][4][4]) {
  unsigned m = 0;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) m |= (a[mt][n][i] > 0.f ? 1u : 0u) << (mt * 16 + n * 4 + i);
  return m;
}
// gate a gradient in C-fragment order by a mask recorded in A-fragment order (i: a1<->c2, a2<->c1)
LP_DEVICE void lp_gate_c(float (&c)[2][4][4], unsigned mask) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const unsigned b = mask >> (mt * 16 + n * 4);
      if (!(b & 1u)) c[mt][n][0] = 0.f;
      if (!(b & 4u)) c[mt][n][1] = 0.f;
      if (!(b & 2u)) c[mt][n][2] = 0.f;
      if (!(b & 8u)) c[mt][n][3] = 0.f;
    }
}
// C-fragment gradient -> TF32 A-fragments of the dX product
LP_DEVICE void lp_c_to_a_tf32(const float (&c)[2][4][4], float (&a)[2][4][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      a[mt][n][0] = lp_tf32_rna(c[mt][n][0]);
      a[mt][n][1] = lp_tf32_rna(c[mt][n][2]);
      a[mt][n][2] = lp_tf32_rna(c[mt][n][1]);
      a[mt][n][3] = lp_tf32_rna(c[mt][n][3]);
    }
}

// dX: acc[mt][nn] += A[mt][jk] * B(jk, nn), B-fragments {hi0, hi1} from the dX weight image
template <int NN, int KS>
LP_DEVICE void lp_dx(const float* Xf, float (&acc)[2][NN][4], const float (&a)[2][KS][4], int lane) {
#pragma unroll
  for (int jk = 0; jk < KS; ++jk)
#pragma unroll
    for (int nn = 0; nn < NN; ++nn) {
      const float2 w = *reinterpret_cast<const float2*>(Xf + ((jk * NN + nn) * 32 + lane) * 2);
      const float b[2] = {w.x, w.y};
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) lp_mma_tf32(acc[mt][nn], a[mt][jk], b);
    }
}

// bf16 operand tiles (byte offsets): feature f (0..31 of the block starting at chunk `chunk0`),
// sample s = 16mt + 8h + g  ->  (chunk0 + f/8)*512 + (2mt+h)*128 + g*16 + (f%8)*2
LP_DEVICE void lp_tile_put_a(unsigned char* stack, int chunk0, const float (&a)[2][4][4], int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 4; ++n) {  // A-fragment order: (a0,a2) = row g, (a1,a3) = row g+8
      unsigned char* p = stack + (chunk0 + n) * LP_TC_SBO + (2 * mt) * LP_TC_LBO + g * 16 + 4 * t;
      *reinterpret_cast<unsigned*>(p) = lp_pack_bf16x2(a[mt][n][0], a[mt][n][2]);
      *reinterpret_cast<unsigned*>(p + LP_TC_LBO) = lp_pack_bf16x2(a[mt][n][1], a[mt][n][3]);
    }
}
LP_DEVICE void lp_tile_put_c(unsigned char* tile, const float (&c)[2][4][4], int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 4; ++n) {  // C-fragment order: (c0,c1) = row g, (c2,c3) = row g+8
      unsigned char* p = tile + n * LP_TC_SBO + (2 * mt) * LP_TC_LBO + g * 16 + 4 * t;
      *reinterpret_cast<unsigned*>(p) = lp_pack_bf16x2(c[mt][n][0], c[mt][n][1]);
      *reinterpret_cast<unsigned*>(p + LP_TC_LBO) = lp_pack_bf16x2(c[mt][n][2], c[mt][n][3]);
    }
}

// per-warp shared memory of the backward kernel (bytes).  A1 stacks the left operands of the four
// 32-wide layers: chunks 0-3 = h1, 4-7 = trunk output, 8.. = grid features x0, then one chunk whose
// first row is all ones (bias gradients).  The tensor core always reads a 16-chunk (M = 128) window;
// the rows behind the used ones are whatever follows in the warp's region and are never read back.
template <int C>
struct BW {
  static constexpr int ONES = 8 + C / 8;          // chunk holding the row of ones = stack row 64 + C
  static constexpr int A1 = 0;
  static constexpr int A2 = A1 + (ONES + 1) * 512;  // colour hidden | opacity hidden | ones : 9 chunks
  static constexpr int DY = A2 + 9 * 512;         // four gradient tiles [32 samples][32]
  static constexpr int DYL = DY + 4 * 2048;       // last-layer gradient tile [32 samples][16]
  static constexpr int RAYS = DYL + 1024;         // fp32 [3][32]: total, g_nlt, g_len per ray
  static constexpr int XS = RAYS + 384;           // fp32 [32][C] transfer tile (ray-owner <-> quad rows)
  static constexpr int END = XS + 32 * C * 4;
  static_assert(A2 + 16 * 512 <= END, "operand window leaves the warp's region");
};
template <int C>
struct BWEnd { static constexpr int value = BW<C>::END; };
// TMEM columns: four A1 x dY_j products, the per-tile encoding product, the last-layer product
constexpr int TM_W = 0, TM_E = 128, TM_L = 160;

// lane 0 of a warp: reduce this warp's 32 samples into the CTA's TMEM accumulators (10 MMAs)
template <int C>
LP_DEVICE void lp_issue_dw(unsigned tmem, unsigned base_lo, const unsigned char* wsb, int accumulate) {
  using B = BW<C>;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      lp_tc_mma_bf16_off(tmem, TM_W + 32 * j, base_lo, wsb, B::A1 + ks * 2 * LP_TC_LBO,
                         B::DY + j * 2048 + ks * 2 * LP_TC_LBO, 32, accumulate | ks);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    lp_tc_mma_bf16_off(tmem, TM_L, base_lo, wsb, B::A2 + ks * 2 * LP_TC_LBO, B::DYL + ks * 2 * LP_TC_LBO, 16,
                       accumulate | ks);
}
// once per ray tile: encoding^T x (sum over steps of the colour-hidden gradient), operands in A1
// chunks 0-3 and gradient tile 2
template <int C>
LP_DEVICE void lp_issue_enc(unsigned tmem, unsigned base_lo, const unsigned char* wsb, int accumulate) {
  using B = BW<C>;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    lp_tc_mma_bf16_off(tmem, TM_E, base_lo, wsb, B::A1 + ks * 2 * LP_TC_LBO, B::DY + 2 * 2048 + ks * 2 * LP_TC_LBO, 32,
                       accumulate | ks);
}

template <int C>
__global__ void __launch_bounds__(256) lp_render_bwd_fast_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G,
                                                                const float* __restrict__ params, LpBwdIo io) {
  using L = Lay<C>;
  using B = BW<C>;
  LP_DYN_SMEM(float, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + L::END);  // [nwarps] + 1
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 9);
  unsigned char* wsb = reinterpret_cast<unsigned char*>(smem + L::END + 32) + warp * BWEnd<C>::value;
  float* wsf = reinterpret_cast<float*>(wsb + B::RAYS);
  float* xs = reinterpret_cast<float*>(wsb + B::XS);

  lp_build_weights<C, true>(smem, params, D);
  for (int e = lane; e < B::RAYS / 4; e += 32) reinterpret_cast<unsigned*>(wsb)[e] = 0u;  // zero all operand tiles
  __syncwarp();
  {  // rows of ones (bf16 1.0 = 0x3F80): A2 stack row 64 (chunk 8) and A1 stack row 64 + C
    const int s = lane;
    *reinterpret_cast<unsigned short*>(wsb + B::A2 + 8 * LP_TC_SBO + (s >> 3) * LP_TC_LBO + (s & 7) * 16) = 0x3F80;
    *reinterpret_cast<unsigned short*>(wsb + B::A1 + B::ONES * LP_TC_SBO + (s >> 3) * LP_TC_LBO + (s & 7) * 16) = 0x3F80;
  }
  const unsigned base_lo = lp_tc_desc_lo(wsb);
  if (threadIdx.x == 0) {
    for (int w = 0; w <= nwarps; ++w) lp_mbar_init(bars + w, 1);
    lp_mbar_init_fence();
  }
  if (warp == 0) lp_tmem_alloc512(tmem_slot);
  lp_fence_async_smem();
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  const unsigned tmem = *tmem_slot;
  if (threadIdx.x == 0) {  // zero the accumulators: D = 0 * 0 with accumulate off (dY tiles are zero)
    lp_issue_dw<C>(tmem, base_lo, wsb, 0);
    lp_issue_enc<C>(tmem, base_lo, wsb, 0);
    lp_tc_commit(bars + nwarps);
  }
  lp_mbar_wait(bars + nwarps, 0);
  lp_tc_fence_after();
  __syncthreads();

  const float* bias = smem + L::BIAS;
  const int num_tiles = (R.n + 31) / 32;
  const int tot = M.S + M.S_inf;
  int iter = 0;  // number of operand hand-offs this warp has made

  const int tiles_per_warp = (num_tiles + gridDim.x * nwarps - 1) / (gridDim.x * nwarps);  // see forward kernel
  for (int it = 0; it < tiles_per_warp; ++it) {
    const int tile = (it * gridDim.x + blockIdx.x) * nwarps + warp;
    const int rbase = tile * 32;
    const Ray1 me = lp_load_ray1(R, rbase + lane, G.g[0].B);
    float rnear[4], rfar[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rnear[i] = __shfl_sync(LP_FULL_MASK, me.near, 8 * i + g);
      rfar[i] = __shfl_sync(LP_FULL_MASK, me.far, 8 * i + g);
    }
    {  // per-ray constants computed by the lane that owns the ray (wsf is private to the warp and not
       // an MMA operand, so no hand-off wait is needed here)
      const int q = me.active ? me.ray : R.n - 1;
      const float gl = me.active ? io.g_len[q] : 0.f, gn = me.active ? io.g_nlt[q] : 0.f;
      float tot_ = gl * io.len[q];
      for (int c = 0; c < D.n_feat; ++c)
        tot_ = fmaf(me.active ? io.g_feat[(long long)q * io.g_feat_stride + c] : 0.f, io.feat[(long long)q * io.feat_stride + c], tot_);
      __syncwarp();
      wsf[lane] = tot_; wsf[32 + lane] = gn; wsf[64 + lane] = gl;
    }
    __syncwarp();
    float nlt[4] = {0.f, 0.f, 0.f, 0.f}, T[4] = {1.f, 1.f, 1.f, 1.f}, prefix[4] = {0.f, 0.f, 0.f, 0.f}, gF[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ray = rbase + 8 * i + g;
      gF[i] = (ray < R.n && t < D.n_feat) ? io.g_feat[(long long)ray * io.g_feat_stride + t] : 0.f;
    }
    float S[2][4][4];  // sum over steps of the colour-hidden gradient (-> encoding gradient, encoding product)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) S[mt][n][i] = 0.f;

    for (int step = 0; step < tot; ++step) {
      if ((step & LP_ALIGN_MASK) == 0) __syncthreads();
      // ------------------------------ forward recompute ------------------------------
      const Sched sc = lp_sched(step, M);
      float depth[4], delta[4];
      unsigned m_h1, m_tr, m_ho, m_hc;
      float last[2][4];
      float sx, sy, sz, soob;  // the owned ray's sample position (reused by the scatter)
      {
        float xa[4][C / 4];
        {
          float dep, del;
          lp_depth_delta(sc, me.near, me.far, dep, del);
          sx = me.ox + dep * me.dx; sy = me.oy + dep * me.dy; sz = me.oz + dep * me.dz;
          if (M.contract) lp_contract(sx, sy, sz);
          soob = M.mask_oob ? lp_in_bounds(sx, sy, sz) : 1.f;
          __syncwarp();
#ifndef LP_ABL_NO_GATHER
          lp_gather_lane<C, true>(G, me.b, sx, sy, sz, soob, xs, lane);
#endif
          __syncwarp();
          // operand tiles of the previous hand-off must have been consumed by the tensor core
          if (iter > 0) lp_mbar_wait(bars + warp, (iter - 1) & 1);
          lp_read_rows<C>(xs, g, t, xa);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          lp_depth_delta(sc, rnear[i], rfar[i], depth[i], delta[i]);
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {  // x0 block of A1: stack rows 64 + 16k + 4t .. +3, sample-group i
            unsigned char* p8 = wsb + B::A1 + (8 + 2 * k + (t >> 1)) * LP_TC_SBO + i * LP_TC_LBO + g * 16 + 8 * (t & 1);
            *reinterpret_cast<uint2*>(p8) = make_uint2(lp_pack_bf16x2(xa[i][4 * k], xa[i][4 * k + 1]),
                                                       lp_pack_bf16x2(xa[i][4 * k + 2], xa[i][4 * k + 3]));
          }
        }
        float acc[2][4][4], a[2][4][4], tr[2][4][4];
        {
          float a0[2][C / 8][4];
          lp_x0_to_a<C>(xa, a0);
          lp_init_bias(acc, bias, t);
          lp_layer3x<C / 8>(smem + L::F_T0, acc, a0, lane);
        }
        lp_relu_to_a(acc, a);
        m_h1 = lp_mask_a(a);
        lp_tile_put_a(wsb + B::A1, 0, a, g, t);
        lp_init_bias(acc, bias + 32, t);
        lp_layer3x<4>(smem + L::F_T1, acc, a, lane);
        lp_relu_to_a(acc, tr);
        m_tr = lp_mask_a(tr);
        lp_tile_put_a(wsb + B::A1, 4, tr, g, t);
        lp_init_bias(acc, bias + 64, t);
        lp_layer3x<4>(smem + L::F_O0, acc, tr, lane);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { last[mt][0] = last[mt][2] = bias[128 + 2 * t]; last[mt][1] = last[mt][3] = bias[128 + 2 * t + 1]; }
        lp_relu_to_a(acc, a);
        m_ho = lp_mask_a(a);
        lp_tile_put_a(wsb + B::A2, 4, a, g, t);
        lp_layer3x_n1(smem + L::F_LO, last, a, lane);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int q0 = min(rbase + 16 * mt + g, R.n - 1), q1 = min(rbase + 16 * mt + 8 + g, R.n - 1);
#ifdef LP_ABL_NO_ENC
            const float2 e0 = make_float2(0.1f * q0, 0.2f), e1 = make_float2(0.3f, 0.1f * q1);
#else
            const float2 e0 = __ldg(reinterpret_cast<const float2*>(R.enc + (long long)q0 * H + 8 * n + 2 * t));
            const float2 e1 = __ldg(reinterpret_cast<const float2*>(R.enc + (long long)q1 * H + 8 * n + 2 * t));
#endif
            tr[mt][n][0] += e0.x; tr[mt][n][2] += e0.y; tr[mt][n][1] += e1.x; tr[mt][n][3] += e1.y;
          }
        lp_init_bias(acc, bias + 96, t);
        lp_layer3x<4>(smem + L::F_C0, acc, tr, lane);
        lp_relu_to_a(acc, a);
        m_hc = lp_mask_a(a);
        lp_tile_put_a(wsb + B::A2, 0, a, g, t);
        lp_layer3x_n1(smem + L::F_LC, last, a, lane);
      }
      // ------------------------------ compositing gradient ------------------------------
      float dl[2][4];  // dY of the last layer in C-fragment order: cols (2t: d_logc_t, 2t+1: g_raw)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int mt = i >> 1, h = i & 1, s = 16 * mt + 8 * h + g;
        float raw = last[mt][2 * h + 1];
        if (M.noise) raw += M.sigma * lp_sample_noise(M, rbase + s, step);
        nlt[i] += delta[i] * M.gain * lp_softplus(raw);
        const float Tn = expf(-nlt[i]);
        const float w = T[i] - Tn;
        T[i] = Tn;
        const float sg = lp_sigmoid(last[mt][2 * h]);
        float pc = (t < D.n_feat) ? sg * gF[i] : 0.f;  // this lane's colour channel
        pc += __shfl_xor_sync(LP_FULL_MASK, pc, 1);
        pc += __shfl_xor_sync(LP_FULL_MASK, pc, 2);
        const float p = fmaf(depth[i], wsf[64 + s], pc);
        prefix[i] = fmaf(w, p, prefix[i]);
        const float suffix = (step == tot - 1) ? 0.f : wsf[s] - prefix[i];
        const float g_dop = Tn * p - suffix + wsf[32 + s];
        dl[mt][2 * h + 1] = g_dop * delta[i] * M.gain * lp_sigmoid(raw);
        dl[mt][2 * h] = (t < D.n_feat) ? w * gF[i] * sg * (1.f - sg) : 0.f;
      }
      // ------------------------------ backward sweep ------------------------------
      float d1[2][4][4], d2[2][4][4], a[2][4][4];
      {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {  // last-layer gradient tile: columns 2t, 2t+1 (chunk 0)
          unsigned char* p = wsb + B::DYL + (2 * mt) * LP_TC_LBO + g * 16 + 4 * t;
          *reinterpret_cast<unsigned*>(p) = lp_pack_bf16x2(dl[mt][0], dl[mt][1]);
          *reinterpret_cast<unsigned*>(p + LP_TC_LBO) = lp_pack_bf16x2(dl[mt][2], dl[mt][3]);
        }
        float al[2][1][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          al[mt][0][0] = lp_tf32_rna(dl[mt][0]); al[mt][0][1] = lp_tf32_rna(dl[mt][2]);
          al[mt][0][2] = lp_tf32_rna(dl[mt][1]); al[mt][0][3] = lp_tf32_rna(dl[mt][3]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) { d1[mt][n][i] = 0.f; d2[mt][n][i] = 0.f; }
        lp_dx<4, 1>(smem + L::X_LC, d1, al, lane);  // d_hc
        lp_dx<4, 1>(smem + L::X_LO, d2, al, lane);  // d_ho
      }
      lp_gate_c(d1, m_hc);
      lp_gate_c(d2, m_ho);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) S[mt][n][i] += d1[mt][n][i];
      lp_tile_put_c(wsb + B::DY + 2 * 2048, d1, g, t);  // dY of the colour hidden layer
      lp_tile_put_c(wsb + B::DY + 1 * 2048, d2, g, t);  // dY of the opacity hidden layer
      lp_c_to_a_tf32(d1, a);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) d1[mt][n][i] = 0.f;
      lp_dx<4, 4>(smem + L::X_C0, d1, a, lane);  // d_xc (C-fragment order)
      lp_c_to_a_tf32(d2, a);
      lp_dx<4, 4>(smem + L::X_O0, d1, a, lane);  // d1 = d_t = d_xc + d_ho' Wo0^T
      lp_gate_c(d1, m_tr);
      lp_tile_put_c(wsb + B::DY + 0 * 2048, d1, g, t);  // dY of trunk layer 1
      lp_c_to_a_tf32(d1, a);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) d2[mt][n][i] = 0.f;
      lp_dx<4, 4>(smem + L::X_T1, d2, a, lane);  // d_h1
      lp_gate_c(d2, m_h1);
      lp_tile_put_c(wsb + B::DY + 3 * 2048, d2, g, t);  // dY of trunk layer 0
      // hand the operand tiles to the tensor core
      lp_fence_async_smem();
      __syncwarp();
#ifndef LP_ABL_NO_DW
      if (lane == 0) {
        lp_tc_fence_after();
        lp_issue_dw<C>(tmem, base_lo, wsb, 1);
        lp_tc_commit(bars + warp);
      }
      ++iter;
#endif
      // input gradient of the first layer and its scatter into the grid
      lp_c_to_a_tf32(d2, a);
      float dx0[2][C / 8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < C / 8; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) dx0[mt][n][i] = 0.f;
      lp_dx<C / 8, 4>(smem + L::X_T0, dx0, a, lane);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // row i = (mt,h) holds channels 16k+4t..+3 in dx0[mt][2k..2k+1][2h..2h+1]
        const int mt = i >> 1, h = i & 1;
#pragma unroll
        for (int k = 0; k < C / 16; ++k)
          *reinterpret_cast<float4*>(xs + lp_xs<C>(8 * i + g, 4 * k + t)) =
              make_float4(dx0[mt][2 * k][2 * h], dx0[mt][2 * k][2 * h + 1], dx0[mt][2 * k + 1][2 * h], dx0[mt][2 * k + 1][2 * h + 1]);
      }
      __syncwarp();
#ifndef LP_ABL_NO_SPLAT
      if (me.active) lp_splat_lane<C>(G, io.g_grid, me.b, sx, sy, sz, soob, xs, lane);
#endif
    }
    // ---- per-tile tail: the ray encoding enters the colour branch as a per-ray constant, so both its
    // gradient (S Wc0^T) and its share of dWc0 (enc^T S) need only the step-sum S ----
    if (iter > 0) lp_mbar_wait(bars + warp, (iter - 1) & 1);
    {
      const int q = min(rbase + lane, R.n - 1);
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)q * H);
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // features 8c..8c+7 of sample `lane` = 16 contiguous bytes of chunk c
        const float4 u = __ldg(e4 + 2 * c), v = __ldg(e4 + 2 * c + 1);
        *reinterpret_cast<uint4*>(wsb + B::A1 + c * LP_TC_SBO + (lane >> 3) * LP_TC_LBO + (lane & 7) * 16) =
            make_uint4(lp_pack_bf16x2(u.x, u.y), lp_pack_bf16x2(u.z, u.w), lp_pack_bf16x2(v.x, v.y), lp_pack_bf16x2(v.z, v.w));
      }
    }
    lp_tile_put_c(wsb + B::DY + 2 * 2048, S, g, t);
    lp_fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      lp_tc_fence_after();
      lp_issue_enc<C>(tmem, base_lo, wsb, 1);
      lp_tc_commit(bars + warp);
    }
    ++iter;
    {
      float a[2][4][4], genc[2][4][4];
      lp_c_to_a_tf32(S, a);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) genc[mt][n][i] = 0.f;
      lp_dx<4, 4>(smem + L::X_C0, genc, a, lane);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int r0 = rbase + 16 * mt + g, r1 = r0 + 8;
          if (r0 < R.n)
            *reinterpret_cast<float2*>(io.g_enc + (long long)r0 * H + 8 * n + 2 * t) = make_float2(genc[mt][n][0], genc[mt][n][1]);
          if (r1 < R.n)
            *reinterpret_cast<float2*>(io.g_enc + (long long)r1 * H + 8 * n + 2 * t) = make_float2(genc[mt][n][2], genc[mt][n][3]);
        }
    }
  }
  // ---- drain: every warp waits for its last hand-off, then the CTA reads the accumulators ----
  if (iter > 0) lp_mbar_wait(bars + warp, (iter - 1) & 1);
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  {
    const LpLayer &t0 = D.trunk.l[0], &t1 = D.trunk.l[1], &o0 = D.opacity.l[0], &o1 = D.opacity.l[1],
                  &c0 = D.color.l[0], &c1 = D.color.l[1];
    float v[32];
    auto add_rows = [&](const LpLayer& Ly, int col) {  // this lane's stack row of a 32-column product
      lp_tmem_ld32(tmem, 32 * warp, col, v);
#pragma unroll
      for (int n = 0; n < 32; ++n) lp_red_add1(io.g_params + Ly.w_off + lane * Ly.N + n, v[n]);
    };
    if (warp == 0) {         // stack rows 0..31: h1 (x d_t), and the encoding product
      add_rows(t1, TM_W + 0);
      add_rows(c0, TM_E);
    } else if (warp == 1) {  // rows 32..63: trunk output (x d_ho, x d_hc)
      add_rows(o0, TM_W + 32);
      add_rows(c0, TM_W + 64);
    } else if (warp == 2) {  // rows 64..64+C: grid features (x d_h1)
      lp_tmem_ld32(tmem, 64, TM_W + 96, v);
      if (lane < C)
        for (int n = 0; n < 32; ++n) lp_red_add1(io.g_params + t0.w_off + lane * t0.N + n, v[n]);
    }
    // bias gradients of the four 32-wide layers = the row of ones (stack row 64 + C) times dY_j
    constexpr int ones_warp = (64 + C) / 32, ones_lane = (64 + C) % 32;
    if (warp == ones_warp) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lp_tmem_ld32(tmem, 32 * ones_warp, TM_W + 32 * j, v);
        const LpLayer& Lb = j == 0 ? t1 : (j == 1 ? o0 : (j == 2 ? c0 : t0));
        if (lane == ones_lane)
          for (int n = 0; n < 32; ++n) lp_red_add1(io.g_params + Lb.b_off + n, v[n]);
      }
    }
    // last layer (A2 x dY_last): rows 0..31 = colour hidden, 32..63 = opacity hidden, 64 = ones
    if (warp < 3) {
      lp_tmem_ld32(tmem, 32 * warp, TM_L, v);
      if (warp == 0) {
        for (int c = 0; c < D.n_feat; ++c) lp_red_add1(io.g_params + c1.w_off + lane * c1.N + c, v[2 * c]);
      } else if (warp == 1) {
        lp_red_add1(io.g_params + o1.w_off + lane, v[1]);
      } else if (lane == 0) {
        for (int c = 0; c < D.n_feat; ++c) lp_red_add1(io.g_params + c1.b_off + c, v[2 * c]);
        lp_red_add1(io.g_params + o1.b_off, v[1]);
      }
    }
  }
  lp_tc_fence_before();
  __syncthreads();
  if (warp == 0) lp_tmem_dealloc512(tmem);
}

}  // namespace lpf

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
static inline bool lp_fast_render_supported(const LpRenderArgs& a) {
  const LpDecoder& D = a.D;
  if (D.use_color_grid || a.use_scaffold) return false;
  if (D.trunk.n_layers != 2 || D.opacity.n_layers != 2 || D.color.n_layers != 2) return false;
  if (D.C != 16 && D.C != 32) return false;
  if (D.n_feat > 3 || D.in_c != lpf::H) return false;
  const LpLayer* ls[4] = {&D.trunk.l[0], &D.trunk.l[1], &D.opacity.l[0], &D.color.l[0]};
  for (int i = 0; i < 4; ++i)
    if (ls[i]->N != lpf::H) return false;
  return true;
}

#ifdef LP_HOSTSIM
#define LP_FAST_SET_SMEM(kernel, bytes) 0
static inline int lp_fast_num_sms() { return 2; }
#else
#define LP_FAST_SET_SMEM(kernel, bytes) \
  (cudaFuncSetAttribute((kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)) != cudaSuccess)
static inline int lp_fast_num_sms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
#endif

template <int C>
static int lp_fast_render_forward_t(cudaStream_t st, const LpRenderArgs& a, const float* params, float* out_len,
                                    float* out_nlt, float* out_feat, int feat_stride) {
  const int warps = 8;
  const size_t bytes = 4ull * (lpf::Lay<C>::FWD_END + warps * (32 * 40 + 32 * C));
  if (LP_FAST_SET_SMEM(lpf::lp_render_fwd_fast_kernel<C>, bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + 31) / 32;
  int blocks = (tiles + warps - 1) / warps;
  const int max_blocks = lp_fast_num_sms() * 2;  // persistent: CTAs loop over ray tiles
  if (blocks > max_blocks) blocks = max_blocks;
  LP_LAUNCH(lpf::lp_render_fwd_fast_kernel<C>, dim3(blocks), dim3(warps * 32), bytes, st, a.R, a.M, a.D, a.G,
            params, out_len, out_nlt, out_feat, feat_stride);
  return LP_OK;
}

static inline int lp_fast_render_forward(cudaStream_t st, const LpRenderArgs& a, const float* params,
                                         float* out_len, float* out_nlt, float* out_feat, int feat_stride) {
  if (a.D.C == 16) return lp_fast_render_forward_t<16>(st, a, params, out_len, out_nlt, out_feat, feat_stride);
  return lp_fast_render_forward_t<32>(st, a, params, out_len, out_nlt, out_feat, feat_stride);
}

static inline bool lp_fast_render_backward_supported(const LpRenderArgs& a) { return a.D.C == 16 || a.D.C == 32; }

template <int C>
static int lp_fast_render_backward_t(cudaStream_t st, const LpRenderArgs& a, const float* params, const LpBwdIo& io) {
  // shared memory: weight image + (barriers, TMEM slot) + 22.4 KB of operand tiles per warp
  const int warps = (C == 16) ? 8 : 7;
  const size_t bytes = 4ull * (lpf::Lay<C>::END + 32) + (size_t)warps * lpf::BWEnd<C>::value;
  if (LP_FAST_SET_SMEM(lpf::lp_render_bwd_fast_kernel<C>, bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + 31) / 32;
  int blocks = (tiles + warps - 1) / warps;
  const int max_blocks = lp_fast_num_sms();  // persistent, one CTA per SM (it owns the SM's TMEM)
  if (blocks > max_blocks) blocks = max_blocks;
  LP_LAUNCH(lpf::lp_render_bwd_fast_kernel<C>, dim3(blocks), dim3(warps * 32), bytes, st, a.R, a.M, a.D, a.G, params, io);
  return LP_OK;
}

static inline int lp_fast_render_backward(cudaStream_t st, const LpRenderArgs& a, const float* params, const LpBwdIo& io) {
  if (a.D.C == 16) return lp_fast_render_backward_t<16>(st, a, params, io);
  return lp_fast_render_backward_t<32>(st, a, params, io);
}
