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
template <int KSTEPS>
LP_DEVICE void lp_layer3x(const float* Wf, float (&acc)[2][4][4], const float (&ain)[2][KSTEPS][4], int lane) {
#pragma unroll
  for (int j = 0; j < KSTEPS; ++j) {
    float ahi[2][4], alo[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) { ahi[mt][i] = lp_tf32_rna(ain[mt][j][i]); alo[mt][i] = ain[mt][j][i] - ahi[mt][i]; }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float4 w = *reinterpret_cast<const float4*>(Wf + ((j * 4 + n) * 32 + lane) * 4);
      const float bh[2] = {w.x, w.y}, bl[2] = {w.z, w.w};
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        lp_mma_tf32(acc[mt][n], alo[mt], bh);
        lp_mma_tf32(acc[mt][n], ahi[mt], bl);
        lp_mma_tf32(acc[mt][n], ahi[mt], bh);
      }
    }
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

// Per-warp ray tile: geometry in registers for the lane's 4 rows (row i = 16*(i>>1) + 8*(i&1) + g).
struct Rows {
  float ox[4], oy[4], oz[4], dx[4], dy[4], dz[4], near[4], far[4];
  int b[4], ray[4];
  bool active[4];
};

LP_DEVICE void lp_load_rows(const LpRays& R, int rbase, int g, int batch, Rows& r) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ray = rbase + 16 * (i >> 1) + 8 * (i & 1) + g;
    r.active[i] = ray < R.n;
    r.ray[i] = ray;
    const int q = r.active[i] ? ray : R.n - 1;
    r.ox[i] = R.org[3 * q]; r.oy[i] = R.org[3 * q + 1]; r.oz[i] = R.org[3 * q + 2];
    r.dx[i] = R.dir[3 * q]; r.dy[i] = R.dir[3 * q + 1]; r.dz[i] = R.dir[3 * q + 2];
    r.near[i] = R.near[q]; r.far[i] = R.far[q];
    r.b[i] = min(max(R.gidx[q], 0), batch - 1);
  }
}

// Gather the lane's channel chunk(s) of one row's sample: xa[C/4] (chunk k = channels 16k+4t..+3).
template <int C>
LP_DEVICE void lp_gather_row(const LpGridSet& G, int b, float x, float y, float z, float oob, int t,
                             float (&xa)[C / 4]) {
#pragma unroll
  for (int k = 0; k < C / 4; ++k) xa[k] = 0.f;
  for (int gi = 0; gi < G.n; ++gi) {
    long long off[8];
    float w[8];
    const int nt = lp_taps(G.g[gi], C, b, x, y, z, off, w);
#pragma unroll
    for (int tp = 0; tp < 8; ++tp) {
      if (tp < nt && w[tp] != 0.f) {
#pragma unroll
        for (int k = 0; k < C / 16; ++k) {
          const float4 v = lp_ldg4(G.data + off[tp] + 16 * k + 4 * t);
          xa[4 * k + 0] = fmaf(w[tp], v.x, xa[4 * k + 0]); xa[4 * k + 1] = fmaf(w[tp], v.y, xa[4 * k + 1]);
          xa[4 * k + 2] = fmaf(w[tp], v.z, xa[4 * k + 2]); xa[4 * k + 3] = fmaf(w[tp], v.w, xa[4 * k + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < C / 4; ++k) xa[k] *= oob;
}

// Sample position, step length etc. of one row at `step`.
struct SamplePos { float x, y, z, depth, delta, oob; };
LP_DEVICE SamplePos lp_sample_pos(const Rows& r, int i, int step, const LpMarch& M) {
  SamplePos p;
  p.depth = lp_depth(step, r.near[i], r.far[i], M.S, M.S_inf, M.disparity_at_inf);
  p.delta = p.depth - lp_depth(step - 1, r.near[i], r.far[i], M.S, M.S_inf, M.disparity_at_inf);
  p.x = r.ox[i] + p.depth * r.dx[i]; p.y = r.oy[i] + p.depth * r.dy[i]; p.z = r.oz[i] + p.depth * r.dz[i];
  if (M.contract) lp_contract(p.x, p.y, p.z);
  p.oob = M.mask_oob ? lp_in_bounds(p.x, p.y, p.z) : 1.f;
  return p;
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
__global__ void __launch_bounds__(256) lp_render_fwd_fast_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G,
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
  float* enc_s = smem + L::FWD_END + warp * (32 * 40);  // per-warp [32 rays][40] encoding tile
  const float* bias = smem + L::BIAS;
  const int num_tiles = (R.n + 31) / 32;
  const int tot = M.S + M.S_inf;

  for (int tile = blockIdx.x * nwarps + warp; tile < num_tiles; tile += gridDim.x * nwarps) {
    const int rbase = tile * 32;
    Rows r;
    lp_load_rows(R, rbase, g, G.g[0].B, r);
    __syncwarp();
    for (int e = lane; e < 32 * 8; e += 32) {  // 32 rays x 8 float4
      const int row = e >> 3, c4 = e & 7;
      const int q = min(rbase + row, R.n - 1);
      *reinterpret_cast<float4*>(enc_s + row * 40 + 4 * c4) = lp_ldg4(R.enc + (long long)q * H + 4 * c4);
    }
    __syncwarp();
    float nlt[4] = {0.f, 0.f, 0.f, 0.f}, T[4] = {1.f, 1.f, 1.f, 1.f}, accum[4] = {0.f, 0.f, 0.f, 0.f};

    for (int step = 0; step < tot; ++step) {
      float xa[4][C / 4], depth[4], delta[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const SamplePos p = lp_sample_pos(r, i, step, M);
        depth[i] = p.depth; delta[i] = p.delta;
        lp_gather_row<C>(G, r.b[i], p.x, p.y, p.z, p.oob, t, xa[i]);
      }
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
        if (M.noise) raw += M.sigma * lp_sample_noise(M, r.ray[i], step);
        nlt[i] += delta[i] * M.gain * lp_softplus(raw);
        const float Tn = expf(-nlt[i]);
        const float w = T[i] - Tn;
        T[i] = Tn;
        accum[i] = fmaf(w, (t == 3) ? depth[i] : lp_sigmoid(last[mt][2 * h]), accum[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (r.active[i]) {
        if (t < D.n_feat) out_feat[(long long)r.ray[i] * feat_stride + t] = accum[i];
        if (t == 3) { out_len[r.ray[i]] = accum[i]; out_nlt[r.ray[i]] = nlt[i]; }
      }
    }
  }
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
  const size_t bytes = 4ull * (lpf::Lay<C>::FWD_END + warps * 32 * 40);
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

static inline bool lp_fast_render_backward_supported(const LpRenderArgs&) { return false; }
static inline int lp_fast_render_backward(cudaStream_t, const LpRenderArgs&, const float*, const LpBwdIo&) {
  return LP_ERR_UNSUPPORTED;
}
