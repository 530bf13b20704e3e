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


// ===========================================================================================
// backward
// ===========================================================================================
// Shared-memory tiles of one warp (32 samples), all [sample][feature] with an XOR swizzle of the
// feature index by 8*(sample&3) so that both the C-fragment stores (float2 per lane) and the
// transposed fragment loads of the dW products are bank-conflict free without padding.
LP_DEVICE int lp_sw32(int s, int f) { return s * 32 + (f ^ ((s & 3) << 3)); }
LP_DEVICE int lp_sw16(int s, int f) { return s * 16 + (f ^ (((s >> 1) & 1) << 3)); }

// store an activation given in A-fragment order (see lp_relu_to_a) as TF32 into a 32-wide tile
LP_DEVICE void lp_store_tile_a(float* tile, const float (&a)[2][4][4], int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      *reinterpret_cast<float2*>(tile + lp_sw32(16 * mt + g, 8 * n + 2 * t)) =
          make_float2(lp_tf32_rna(a[mt][n][0]), lp_tf32_rna(a[mt][n][2]));
      *reinterpret_cast<float2*>(tile + lp_sw32(16 * mt + 8 + g, 8 * n + 2 * t)) =
          make_float2(lp_tf32_rna(a[mt][n][1]), lp_tf32_rna(a[mt][n][3]));
    }
}
// same for a gradient held in C-fragment order
LP_DEVICE void lp_store_tile_c(float* tile, const float (&c)[2][4][4], int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      *reinterpret_cast<float2*>(tile + lp_sw32(16 * mt + g, 8 * n + 2 * t)) =
          make_float2(lp_tf32_rna(c[mt][n][0]), lp_tf32_rna(c[mt][n][1]));
      *reinterpret_cast<float2*>(tile + lp_sw32(16 * mt + 8 + g, 8 * n + 2 * t)) =
          make_float2(lp_tf32_rna(c[mt][n][2]), lp_tf32_rna(c[mt][n][3]));
    }
}

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

// dW tile update over the warp's 32 samples:  acc[mf][nn] += X^T(features 16mf.., samples) * dY
// X tile [32][KF] (swizzled), dY tile [32][NO*8]; accumulators live in smem in fragment order.
template <int KF, int NO>
LP_DEVICE void lp_dw(float* accum, const float* X, const float* dY, int lane) {
  const int g = lane >> 2, t = lane & 3;
  constexpr int MF = KF / 16;
  float c[MF][NO][4];
#pragma unroll
  for (int mf = 0; mf < MF; ++mf)
#pragma unroll
    for (int nn = 0; nn < NO; ++nn) {
      const float4 v = *reinterpret_cast<const float4*>(accum + ((mf * NO + nn) * 32 + lane) * 4);
      c[mf][nn][0] = v.x; c[mf][nn][1] = v.y; c[mf][nn][2] = v.z; c[mf][nn][3] = v.w;
    }
#pragma unroll
  for (int js = 0; js < 4; ++js) {
    float a[MF][4];
#pragma unroll
    for (int mf = 0; mf < MF; ++mf) {
      if (KF == 16) {
        a[mf][0] = X[lp_sw16(8 * js + t, g)];     a[mf][1] = X[lp_sw16(8 * js + t, g + 8)];
        a[mf][2] = X[lp_sw16(8 * js + t + 4, g)]; a[mf][3] = X[lp_sw16(8 * js + t + 4, g + 8)];
      } else {
        a[mf][0] = X[lp_sw32(8 * js + t, 16 * mf + g)];     a[mf][1] = X[lp_sw32(8 * js + t, 16 * mf + g + 8)];
        a[mf][2] = X[lp_sw32(8 * js + t + 4, 16 * mf + g)]; a[mf][3] = X[lp_sw32(8 * js + t + 4, 16 * mf + g + 8)];
      }
    }
#pragma unroll
    for (int nn = 0; nn < NO; ++nn) {
      const float b[2] = {dY[lp_sw32(8 * js + t, 8 * nn + g)], dY[lp_sw32(8 * js + t + 4, 8 * nn + g)]};
#pragma unroll
      for (int mf = 0; mf < MF; ++mf) lp_mma_tf32(c[mf][nn], a[mf], b);
    }
  }
#pragma unroll
  for (int mf = 0; mf < MF; ++mf)
#pragma unroll
    for (int nn = 0; nn < NO; ++nn)
      *reinterpret_cast<float4*>(accum + ((mf * NO + nn) * 32 + lane) * 4) =
          make_float4(c[mf][nn][0], c[mf][nn][1], c[mf][nn][2], c[mf][nn][3]);
}

// bias gradient: lane n accumulates the column sum of the dY tile
LP_DEVICE float lp_colsum(const float* dY, int lane) {
  float s = 0.f;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) s += dY[lp_sw32(r, lane)];
  return s;
}

// scatter one row's channel chunk(s) of d_x0 into the grid gradient
template <int C>
LP_DEVICE void lp_splat_row(const LpGridSet& G, float* grad, int b, float x, float y, float z, float oob, int t,
                            const float (&d)[C / 4]) {
  if (oob == 0.f) return;
  for (int gi = 0; gi < G.n; ++gi) {
    long long off[8];
    float w[8];
    const int nt = lp_taps(G.g[gi], C, b, x, y, z, off, w);
#pragma unroll
    for (int tp = 0; tp < 8; ++tp) {
      if (tp < nt && w[tp] != 0.f) {
        const float ww = w[tp] * oob;
#pragma unroll
        for (int k = 0; k < C / 16; ++k)
          lp_red_add4(grad + off[tp] + 16 * k + 4 * t, ww * d[4 * k], ww * d[4 * k + 1], ww * d[4 * k + 2],
                      ww * d[4 * k + 3]);
      }
    }
  }
}

// per-warp shared-memory map of the backward kernel (floats)
template <int C>
struct BLay {
  static constexpr int X0 = 0;                 // [32][C]
  static constexpr int H1 = X0 + 32 * C;       // [32][32] each
  static constexpr int TR = H1 + 1024;
  static constexpr int XC = TR + 1024;
  static constexpr int HO = XC + 1024;
  static constexpr int HC = HO + 1024;
  static constexpr int DY = HC + 1024;         // gradient tile (B operand of dW, bias sums)
  static constexpr int AW_T0 = DY + 1024;      // dW accumulators, fragment order [mf][nn][32][4]
  static constexpr int AW_T1 = AW_T0 + C * 32;
  static constexpr int AW_O0 = AW_T1 + 1024;
  static constexpr int AW_C0 = AW_O0 + 1024;
  static constexpr int AW_LC = AW_C0 + 1024;   // hc^T * dY_last [2][1][32][4]
  static constexpr int AW_LO = AW_LC + 256;    // ho^T * dY_last
  static constexpr int RAYS = AW_LO + 256;     // per ray: total, g_nlt, g_len  [3][32]
  static constexpr int END = RAYS + 96;
};

template <int C>
__global__ void __launch_bounds__(128) lp_render_bwd_fast_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G,
                                                                const float* __restrict__ params, LpBwdIo io) {
  using L = Lay<C>;
  using B = BLay<C>;
  LP_DYN_SMEM(float, smem);
  lp_build_weights<C, true>(smem, params, D);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* ws = smem + L::END + warp * B::END;
  for (int e = lane; e < B::RAYS - B::AW_T0; e += 32) ws[B::AW_T0 + e] = 0.f;
  float db[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // bias gradients: lane n owns column n (t0,t1,o0,c0,last)
  __syncthreads();
  const float* bias = smem + L::BIAS;
  const int num_tiles = (R.n + 31) / 32;
  const int tot = M.S + M.S_inf;

  for (int tile = blockIdx.x * nwarps + warp; tile < num_tiles; tile += gridDim.x * nwarps) {
    const int rbase = tile * 32;
    Rows r;
    lp_load_rows(R, rbase, g, G.g[0].B, r);
    __syncwarp();
    {  // per-ray constants computed by the lane that owns the ray
      const int ray = rbase + lane;
      const bool act = ray < R.n;
      const int q = act ? ray : R.n - 1;
      float gl = act ? io.g_len[q] : 0.f, gn = act ? io.g_nlt[q] : 0.f;
      float tot_ = gl * io.len[q];
      for (int c = 0; c < D.n_feat; ++c)
        tot_ = fmaf(act ? io.g_feat[(long long)q * io.g_feat_stride + c] : 0.f, io.feat[(long long)q * io.feat_stride + c], tot_);
      ws[B::RAYS + lane] = tot_; ws[B::RAYS + 32 + lane] = gn; ws[B::RAYS + 64 + lane] = gl;
    }
    __syncwarp();
    float nlt[4] = {0.f, 0.f, 0.f, 0.f}, T[4] = {1.f, 1.f, 1.f, 1.f}, prefix[4] = {0.f, 0.f, 0.f, 0.f}, gF[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      gF[i] = (r.active[i] && t < D.n_feat) ? io.g_feat[(long long)r.ray[i] * io.g_feat_stride + t] : 0.f;
    float genc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) genc[mt][n][i] = 0.f;

    for (int step = 0; step < tot; ++step) {
      // ------------------------------ forward recompute ------------------------------
      float depth[4], delta[4];
      unsigned m_h1, m_tr, m_ho, m_hc;
      float last[2][4];
      {
        float xa[4][C / 4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const SamplePos p = lp_sample_pos(r, i, step, M);
          depth[i] = p.depth; delta[i] = p.delta;
          lp_gather_row<C>(G, r.b[i], p.x, p.y, p.z, p.oob, t, xa[i]);
          const int s = 16 * (i >> 1) + 8 * (i & 1) + g;
#pragma unroll
          for (int k = 0; k < C / 16; ++k) {
            const float4 v = make_float4(lp_tf32_rna(xa[i][4 * k]), lp_tf32_rna(xa[i][4 * k + 1]),
                                         lp_tf32_rna(xa[i][4 * k + 2]), lp_tf32_rna(xa[i][4 * k + 3]));
            if (C == 16) *reinterpret_cast<float4*>(ws + B::X0 + lp_sw16(s, 4 * t)) = v;
            else *reinterpret_cast<float4*>(ws + B::X0 + lp_sw32(s, 16 * k + 4 * t)) = v;
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
        lp_store_tile_a(ws + B::H1, a, g, t);
        lp_init_bias(acc, bias + 32, t);
        lp_layer3x<4>(smem + L::F_T1, acc, a, lane);
        lp_relu_to_a(acc, tr);
        m_tr = lp_mask_a(tr);
        lp_store_tile_a(ws + B::TR, tr, g, t);
        lp_init_bias(acc, bias + 64, t);
        lp_layer3x<4>(smem + L::F_O0, acc, tr, lane);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { last[mt][0] = last[mt][2] = bias[128 + 2 * t]; last[mt][1] = last[mt][3] = bias[128 + 2 * t + 1]; }
        lp_relu_to_a(acc, a);
        m_ho = lp_mask_a(a);
        lp_store_tile_a(ws + B::HO, a, g, t);
        lp_layer3x_n1(smem + L::F_LO, last, a, lane);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int q0 = r.active[2 * mt] ? r.ray[2 * mt] : R.n - 1, q1 = r.active[2 * mt + 1] ? r.ray[2 * mt + 1] : R.n - 1;
            const float2 e0 = __ldg(reinterpret_cast<const float2*>(R.enc + (long long)q0 * H + 8 * n + 2 * t));
            const float2 e1 = __ldg(reinterpret_cast<const float2*>(R.enc + (long long)q1 * H + 8 * n + 2 * t));
            tr[mt][n][0] += e0.x; tr[mt][n][2] += e0.y; tr[mt][n][1] += e1.x; tr[mt][n][3] += e1.y;
          }
        lp_store_tile_a(ws + B::XC, tr, g, t);
        lp_init_bias(acc, bias + 96, t);
        lp_layer3x<4>(smem + L::F_C0, acc, tr, lane);
        lp_relu_to_a(acc, a);
        m_hc = lp_mask_a(a);
        lp_store_tile_a(ws + B::HC, a, g, t);
        lp_layer3x_n1(smem + L::F_LC, last, a, lane);
      }
      // ------------------------------ compositing gradient ------------------------------
      float dl[2][4];  // dY of the last layer in C-fragment order: cols (2t: d_logc_t, 2t+1: g_raw)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int mt = i >> 1, h = i & 1, s = 16 * mt + 8 * h + g;
        float raw = last[mt][2 * h + 1];
        if (M.noise) raw += M.sigma * lp_sample_noise(M, r.ray[i], step);
        nlt[i] += delta[i] * M.gain * lp_softplus(raw);
        const float Tn = expf(-nlt[i]);
        const float w = T[i] - Tn;
        T[i] = Tn;
        const float sg = lp_sigmoid(last[mt][2 * h]);
        float pc = (t < D.n_feat) ? sg * gF[i] : 0.f;  // this lane's colour channel
        pc += __shfl_xor_sync(LP_FULL_MASK, pc, 1);
        pc += __shfl_xor_sync(LP_FULL_MASK, pc, 2);
        const float p = fmaf(depth[i], ws[B::RAYS + 64 + s], pc);
        prefix[i] = fmaf(w, p, prefix[i]);
        const float suffix = (step == tot - 1) ? 0.f : ws[B::RAYS + s] - prefix[i];
        const float g_dop = Tn * p - suffix + ws[B::RAYS + 32 + s];
        dl[mt][2 * h + 1] = g_dop * delta[i] * M.gain * lp_sigmoid(raw);
        dl[mt][2 * h] = (t < D.n_feat) ? w * gF[i] * sg * (1.f - sg) : 0.f;
      }
      // ------------------------------ backward sweep ------------------------------
      float d1[2][4][4], d2[2][4][4], a[2][4][4];
      // last layer: dW (hc^T dY, ho^T dY), db, and d_hc / d_ho
      {
        __syncwarp();
        // dY_last tile: 8 valid columns (cols 8..31 of the tile are not read by the NO=1 product)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          *reinterpret_cast<float2*>(ws + B::DY + lp_sw32(16 * mt + g, 2 * t)) = make_float2(lp_tf32_rna(dl[mt][0]), lp_tf32_rna(dl[mt][1]));
          *reinterpret_cast<float2*>(ws + B::DY + lp_sw32(16 * mt + 8 + g, 2 * t)) = make_float2(lp_tf32_rna(dl[mt][2]), lp_tf32_rna(dl[mt][3]));
        }
        __syncwarp();
        lp_dw<32, 1>(ws + B::AW_LC, ws + B::HC, ws + B::DY, lane);
        lp_dw<32, 1>(ws + B::AW_LO, ws + B::HO, ws + B::DY, lane);
        if (lane < 8) {
          float sacc = 0.f;
          for (int rr = 0; rr < 32; ++rr) sacc += ws[B::DY + lp_sw32(rr, lane)];
          db[4] += sacc;
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
      // colour hidden layer: dW_c0 += xc^T d_hc', d_xc = d_hc' Wc0^T
      __syncwarp();
      lp_store_tile_c(ws + B::DY, d1, g, t);
      __syncwarp();
      lp_dw<32, 4>(ws + B::AW_C0, ws + B::XC, ws + B::DY, lane);
      db[3] += lp_colsum(ws + B::DY, lane);
      lp_c_to_a_tf32(d1, a);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) d1[mt][n][i] = 0.f;
      lp_dx<4, 4>(smem + L::X_C0, d1, a, lane);  // d_xc (C-fragment order)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) genc[mt][n][i] += d1[mt][n][i];
      // opacity hidden layer: dW_o0 += t^T d_ho', d_t = d_xc + d_ho' Wo0^T
      __syncwarp();
      lp_store_tile_c(ws + B::DY, d2, g, t);
      __syncwarp();
      lp_dw<32, 4>(ws + B::AW_O0, ws + B::TR, ws + B::DY, lane);
      db[2] += lp_colsum(ws + B::DY, lane);
      lp_c_to_a_tf32(d2, a);
      lp_dx<4, 4>(smem + L::X_O0, d1, a, lane);  // d1 = d_t
      lp_gate_c(d1, m_tr);
      // trunk layer 1
      __syncwarp();
      lp_store_tile_c(ws + B::DY, d1, g, t);
      __syncwarp();
      lp_dw<32, 4>(ws + B::AW_T1, ws + B::H1, ws + B::DY, lane);
      db[1] += lp_colsum(ws + B::DY, lane);
      lp_c_to_a_tf32(d1, a);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) d2[mt][n][i] = 0.f;
      lp_dx<4, 4>(smem + L::X_T1, d2, a, lane);  // d_h1
      lp_gate_c(d2, m_h1);
      // trunk layer 0
      __syncwarp();
      lp_store_tile_c(ws + B::DY, d2, g, t);
      __syncwarp();
      lp_dw<C, 4>(ws + B::AW_T0, ws + B::X0, ws + B::DY, lane);
      db[0] += lp_colsum(ws + B::DY, lane);
      lp_c_to_a_tf32(d2, a);
      float dx0[2][C / 8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int n = 0; n < C / 8; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) dx0[mt][n][i] = 0.f;
      lp_dx<C / 8, 4>(smem + L::X_T0, dx0, a, lane);
      // scatter d_x0: row (mt,h) holds channels 16k+4t..+3 in (dx0[mt][2k][2h..], dx0[mt][2k+1][2h..])
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int mt = i >> 1, h = i & 1;
        float d[C / 4];
#pragma unroll
        for (int k = 0; k < C / 16; ++k) {
          d[4 * k] = dx0[mt][2 * k][2 * h]; d[4 * k + 1] = dx0[mt][2 * k][2 * h + 1];
          d[4 * k + 2] = dx0[mt][2 * k + 1][2 * h]; d[4 * k + 3] = dx0[mt][2 * k + 1][2 * h + 1];
        }
        if (r.active[i]) {
          const SamplePos p = lp_sample_pos(r, i, step, M);
          lp_splat_row<C>(G, io.g_grid, r.b[i], p.x, p.y, p.z, p.oob, t, d);
        }
      }
    }
    // ray-encoding gradient of this tile
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        if (r.active[2 * mt])
          *reinterpret_cast<float2*>(io.g_enc + (long long)r.ray[2 * mt] * H + 8 * n + 2 * t) = make_float2(genc[mt][n][0], genc[mt][n][1]);
        if (r.active[2 * mt + 1])
          *reinterpret_cast<float2*>(io.g_enc + (long long)r.ray[2 * mt + 1] * H + 8 * n + 2 * t) = make_float2(genc[mt][n][2], genc[mt][n][3]);
      }
  }
  // ---- flush the warp's parameter-gradient accumulators ----
  __syncwarp();
  {
    const LpLayer &t0 = D.trunk.l[0], &t1 = D.trunk.l[1], &o0 = D.opacity.l[0], &o1 = D.opacity.l[1],
                  &c0 = D.color.l[0], &c1 = D.color.l[1];
    auto flush = [&](const float* acc, const LpLayer& Ly, int mfs) {
      for (int mf = 0; mf < mfs; ++mf)
        for (int nn = 0; nn < 4; ++nn) {
          const float4 v = *reinterpret_cast<const float4*>(acc + ((mf * 4 + nn) * 32 + lane) * 4);
          float* dst = io.g_params + Ly.w_off;
          lp_red_add1(dst + (16 * mf + g) * Ly.N + 8 * nn + 2 * t, v.x);
          lp_red_add1(dst + (16 * mf + g) * Ly.N + 8 * nn + 2 * t + 1, v.y);
          lp_red_add1(dst + (16 * mf + g + 8) * Ly.N + 8 * nn + 2 * t, v.z);
          lp_red_add1(dst + (16 * mf + g + 8) * Ly.N + 8 * nn + 2 * t + 1, v.w);
        }
    };
    flush(ws + B::AW_T0, t0, C / 16);
    flush(ws + B::AW_T1, t1, 2);
    flush(ws + B::AW_O0, o0, 2);
    flush(ws + B::AW_C0, c0, 2);
    for (int mf = 0; mf < 2; ++mf) {  // last layer: columns (2t: colour t | 2t+1: opacity for t == 0)
      const float4 vc = *reinterpret_cast<const float4*>(ws + B::AW_LC + (mf * 32 + lane) * 4);
      const float4 vo = *reinterpret_cast<const float4*>(ws + B::AW_LO + (mf * 32 + lane) * 4);
      if (t < D.n_feat) {
        lp_red_add1(io.g_params + c1.w_off + (16 * mf + g) * c1.N + t, vc.x);
        lp_red_add1(io.g_params + c1.w_off + (16 * mf + g + 8) * c1.N + t, vc.z);
      }
      if (t == 0) {
        lp_red_add1(io.g_params + o1.w_off + (16 * mf + g), vo.y);
        lp_red_add1(io.g_params + o1.w_off + (16 * mf + g + 8), vo.w);
      }
    }
    lp_red_add1(io.g_params + t0.b_off + lane, db[0]);
    lp_red_add1(io.g_params + t1.b_off + lane, db[1]);
    lp_red_add1(io.g_params + o0.b_off + lane, db[2]);
    lp_red_add1(io.g_params + c0.b_off + lane, db[3]);
    if (lane < 8) {
      if (lane & 1) { if (lane == 1) lp_red_add1(io.g_params + o1.b_off, db[4]); }
      else if ((lane >> 1) < D.n_feat) lp_red_add1(io.g_params + c1.b_off + (lane >> 1), db[4]);
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

static inline bool lp_fast_render_backward_supported(const LpRenderArgs& a) { return a.D.C == 16 || a.D.C == 32; }

template <int C>
static int lp_fast_render_backward_t(cudaStream_t st, const LpRenderArgs& a, const float* params, const LpBwdIo& io) {
  // C=16: 4 warps x 38.4 KB + 49.7 KB weight image; C=32: 3 warps (shared-memory budget)
  const int warps = (C == 16) ? 4 : 3;
  const size_t bytes = 4ull * (lpf::Lay<C>::END + warps * lpf::BLay<C>::END);
  if (LP_FAST_SET_SMEM(lpf::lp_render_bwd_fast_kernel<C>, bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + 31) / 32;
  int blocks = (tiles + warps - 1) / warps;
  const int max_blocks = lp_fast_num_sms();  // persistent, one CTA per SM
  if (blocks > max_blocks) blocks = max_blocks;
  LP_LAUNCH(lpf::lp_render_bwd_fast_kernel<C>, dim3(blocks), dim3(warps * 32), bytes, st, a.R, a.M, a.D, a.G, params, io);
  return LP_OK;
}

static inline int lp_fast_render_backward(cudaStream_t st, const LpRenderArgs& a, const float* params, const LpBwdIo& io) {
  if (a.D.C == 16) return lp_fast_render_backward_t<16>(st, a, params, io);
  return lp_fast_render_backward_t<32>(st, a, params, io);
}
