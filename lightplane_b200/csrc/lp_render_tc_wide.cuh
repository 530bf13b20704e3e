// Tensor-core Renderer kernels for hidden width 64 (the reference's example configuration,
// examples/config/synthetic_overfit.json: mlp_hidden_chn 64, 32-channel triplane, occupancy scaffold): same
// thread-per-sample tcgen05 scheme as lp_render_tc.cuh (read its header first), every per-sample row 64 wide:
//   t0  [C -> 64]   K = C,   N = 64        t1  [64 -> 64]  K = 64,  N = 64
//   o0 | c0         K = 64,  N = 128  (+ the ray-encoding product into the colour half, K = 64, N = 64)
// Tensor memory per group: A hi 0..31 / lo 32..63, encoding hi 64..95 / lo 96..127, D 128..255 -> two groups of 128
// threads per SM.  The backward (second half of this file) runs one group of 256 threads per SM.
#pragma once

#include "lp_render_tc.cuh"

namespace lptc {

constexpr int HW = 64;  // hidden width served here

template <int C>
struct WImg {
  static constexpr int T0_HI = 0;                          // [64 out][C in]
  static constexpr int T0_LO = T0_HI + HW * C * 2;
  static constexpr int T1_HI = T0_LO + HW * C * 2;         // [64][64]
  static constexpr int T1_LO = T1_HI + HW * HW * 2;
  static constexpr int OC_HI = T1_LO + HW * HW * 2;        // [128 out: opacity hidden | colour hidden][64 in]
  static constexpr int OC_LO = OC_HI + 2 * HW * HW * 2;
  static constexpr int F32 = OC_LO + 2 * HW * HW * 2;      // fp32: b_t0 b_t1 b_o0 b_c0 [4][64] | wo1[64] | Wc1[64][4] | b_last[4]
  static constexpr int FB = 0, FWO = 4 * HW, FWC = 5 * HW, FBL = 9 * HW, NF = 9 * HW + 4;
  static constexpr int END = (F32 + NF * 4 + 127) / 128 * 128;
};
constexpr int WT_A = 0, WT_E = 64, WT_D = 128, WT_GROUP_COLS = 256;

template <int C>
LP_DEVICE void lp_build_wimg(unsigned char* sm, const float* __restrict__ P, const LpDecoder& D) {
  using I = WImg<C>;
  const LpLayer &t0 = D.trunk.l[0], &t1 = D.trunk.l[1], &o0 = D.opacity.l[0], &o1 = D.opacity.l[1],
                &c0 = D.color.l[0], &c1 = D.color.l[1];
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int e = tid; e < HW * C; e += nth) lp_put_w(sm, I::T0_HI, I::T0_LO, e % HW, e / HW, C, P[t0.w_off + (e / HW) * t0.N + (e % HW)]);
  for (int e = tid; e < HW * HW; e += nth) {
    const int n = e % HW, k = e / HW;
    lp_put_w(sm, I::T1_HI, I::T1_LO, n, k, HW, P[t1.w_off + k * t1.N + n]);
    lp_put_w(sm, I::OC_HI, I::OC_LO, n, k, HW, P[o0.w_off + k * o0.N + n]);
    lp_put_w(sm, I::OC_HI, I::OC_LO, HW + n, k, HW, P[c0.w_off + k * c0.N + n]);
  }
  float* F = reinterpret_cast<float*>(sm + I::F32);
  for (int e = tid; e < HW; e += nth) {
    F[I::FB + e] = P[t0.b_off + e];
    F[I::FB + HW + e] = P[t1.b_off + e];
    F[I::FB + 2 * HW + e] = P[o0.b_off + e];
    F[I::FB + 3 * HW + e] = P[c0.b_off + e];
    F[I::FWO + e] = P[o1.w_off + e * o1.N];
    for (int c = 0; c < 4; ++c) F[I::FWC + 4 * e + c] = c < D.n_feat ? P[c1.w_off + e * c1.N + c] : 0.f;
  }
  if (tid < 4) F[I::FBL + tid] = tid == 3 ? P[o1.b_off] : (tid < D.n_feat ? P[c1.b_off + tid] : 0.f);
}

// this thread's 64-wide accumulator row: + bias, ReLU
LP_DEVICE void lp_wide_ld_relu(unsigned taddr, const float* bias, float (&v)[HW]) {
  float a[32];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    lp_tmem_ld32u(taddr + 32 * half, a);
    lp_tmem_zero<32>(taddr + 32 * half);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[32 * half + j] = fmaxf(a[j] + bias[32 * half + j], 0.f);
  }
}

template <int C, bool SCAF>
__global__ void __launch_bounds__(256, 1) lp_render_fwd_tcw_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G, LpGridSet SC,
                                                                    const float* __restrict__ params,
                                                                    float* __restrict__ out_len, float* __restrict__ out_nlt,
                                                                    float* __restrict__ out_feat, int feat_stride) {
  using I = WImg<C>;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x;
  const int grp = tid / GT, ngroups = blockDim.x / GT, wig = (tid >> 5) & 3;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + I::END);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  lp_build_wimg<C>(sm, params, D);
  if (tid == 0) {
    for (int i = 0; i < ngroups; ++i) lp_mbar_init(bars + i, 4);
    lp_mbar_init_fence();
  }
  if (tid < 32) lp_tmem_alloc512(tmem_slot);
  lp_fence_async_smem();
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  const unsigned tbase = *tmem_slot + (unsigned)(grp * WT_GROUP_COLS);
  const unsigned tme = lp_taddr(tbase, wig, 0);
  const bool issuer = (tid & 31) == 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) lp_tmem_zero<32>(tme + WT_D + 32 * k);
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  const lp_kdesc_t w_t0h = lp_tc_kdesc_lo(sm + I::T0_HI), w_t0l = lp_tc_kdesc_lo(sm + I::T0_LO),
                   w_t1h = lp_tc_kdesc_lo(sm + I::T1_HI), w_t1l = lp_tc_kdesc_lo(sm + I::T1_LO),
                   w_och = lp_tc_kdesc_lo(sm + I::OC_HI), w_ocl = lp_tc_kdesc_lo(sm + I::OC_LO);
  constexpr int NS = (HW / 8) * 128;  // n-chunk stride of the K = 64 tiles
  unsigned long long* bar = bars + grp;
  int phase = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;

#define LP_W_ROUND(ISSUE) LP_TCG_HANDOFF(1 + grp, GT, lp_elect_one(), ISSUE; lp_tc_commit(bar)) LP_TCG_WAIT(bar, phase)

  for (int tile = blockIdx.x * ngroups + grp; tile < num_tiles; tile += gridDim.x * ngroups) {
    const Ray1 me = lp_load_ray1(R, tile * GT + (tid % GT), G.g[0].B);
    {  // the ray encoding: second half of the colour layer's K = 128, staged once per ray
      float e[HW];
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)(me.active ? me.ray : R.n - 1) * HW);
#pragma unroll
      for (int k = 0; k < HW / 4; ++k) {
        const float4 v = __ldg(e4 + k);
        e[4 * k] = v.x; e[4 * k + 1] = v.y; e[4 * k + 2] = v.z; e[4 * k + 3] = v.w;
      }
      lp_stage_row<HW, 32>(tme + WT_E, e);
    }
    LpCompFwd cf;
    float e_raw = 0.f, e_lg0 = 0.f, e_lg1 = 0.f, e_lg2 = 0.f;  // empty-space folding, see lp_render_fwd_tc_kernel

    for (int step = LP_TC_EMPTY_FOLD ? -1 : 0; step < tot; ++step) {
      const bool probe = step < 0;
      const Sched sc = lp_sched(probe ? 0 : step, M);
      float depth, delta;
      lp_depth_delta(sc, me.near, me.far, depth, delta);
      float occ = 1.f;
      bool hit = false;
      {
        float x0[C];
        if (!probe) {
          float x = me.ox + depth * me.dx, y = me.oy + depth * me.dy, z = me.oz + depth * me.dz;
          if (M.contract) lp_contract(x, y, z);
          if (SCAF) {
            occ = lp_nearest(SC, me.b, x, y, z);
            if (!lp_bar_any(1 + grp, GT, occ != 0.f)) continue;
          }
          const float oob = M.mask_oob ? lp_in_bounds(x, y, z) : 1.f;
          hit = lp_gather_regs<C>(G, me.b, x, y, z, oob, x0);
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) x0[c] = 0.f;
        }
        lp_stage_row<C, 32>(tme + WT_A, x0);
      }
      float raw, lg0, lg1, lg2;
      lp_tmem_wait_st();
      lp_tc_fence_before();
      const bool full = LP_TC_EMPTY_FOLD ? (lp_bar_any(1 + grp, GT, hit) || probe) : (lp_bar_sync(1 + grp, GT), true);
      if (full) {
        if (issuer) {
          lp_tc_fence_after();
          lp_issue_layer_part(tbase, WT_D, WT_A, w_t0h, w_t0l, C / 16, 0, (C / 8) * 128, HW, 32, wig);
          lp_tc_commit(bar);
        }
        lp_mbar_wait(bar, phase); phase ^= 1;
        lp_tc_fence_after();
        float v[HW];
        lp_wide_ld_relu(tme + WT_D, F + I::FB, v);
        lp_stage_row<HW, 32>(tme + WT_A, v);
        LP_W_ROUND(lp_issue_layer_part(tbase, WT_D, WT_A, w_t1h, w_t1l, 4, 0, NS, HW, 32, wig));
        lp_wide_ld_relu(tme + WT_D, F + I::FB + HW, v);
        lp_stage_row<HW, 32>(tme + WT_A, v);
        // opacity | colour hidden layers on the trunk output (N = 128), then the encoding into the colour half
        LP_W_ROUND(lp_issue_layer_part(tbase, WT_D, WT_A, w_och, w_ocl, 4, 0, NS, 2 * HW, 32, wig);
                   lp_issue_layer_part(tbase, WT_D + HW, WT_E, lp_tc_kadv(w_och, (HW / 8) * NS), lp_tc_kadv(w_ocl, (HW / 8) * NS), 4,
                                       0, NS, HW, 32, wig));
        raw = F[I::FBL + 3]; lg0 = F[I::FBL]; lg1 = F[I::FBL + 1]; lg2 = F[I::FBL + 2];
        lp_wide_ld_relu(tme + WT_D, F + I::FB + 2 * HW, v);
#pragma unroll
        for (int j = 0; j < HW; ++j) raw = fmaf(v[j], F[I::FWO + j], raw);
        lp_wide_ld_relu(tme + WT_D + HW, F + I::FB + 3 * HW, v);
#pragma unroll
        for (int j = 0; j < HW; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(F + I::FWC + 4 * j);
          lg0 = fmaf(v[j], w.x, lg0); lg1 = fmaf(v[j], w.y, lg1); lg2 = fmaf(v[j], w.z, lg2);
        }
        if (probe) { e_raw = raw; e_lg0 = lg0; e_lg1 = lg1; e_lg2 = lg2; continue; }
      } else {
        raw = e_raw; lg0 = e_lg0; lg1 = e_lg1; lg2 = e_lg2;
      }
      // ---- compositing (renderer_fw.py:289-340) ----
      cf.add(M, me.ray, step, raw, lg0, lg1, lg2, depth, delta, occ);
    }
    if (me.active) {
      out_len[me.ray] = cf.len;
      out_nlt[me.ray] = cf.nlt;
      for (int c = 0; c < D.n_feat; ++c) out_feat[(long long)me.ray * feat_stride + c] = c == 0 ? cf.c0 : (c == 1 ? cf.c1 : cf.c2);
    }
  }
#undef LP_W_ROUND
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(*tmem_slot);
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
static inline bool lp_tcw_forward_supported(const LpRenderArgs& a) {
  const LpDecoder& D = a.D;
  if (D.use_color_grid) return false;
  if (D.trunk.n_layers != 2 || D.opacity.n_layers != 2 || D.color.n_layers != 2) return false;
  if (D.C != 16 && D.C != 32) return false;
  if (D.n_feat > 3 || D.in_c != HW) return false;
  const LpLayer* ls[4] = {&D.trunk.l[0], &D.trunk.l[1], &D.opacity.l[0], &D.color.l[0]};
  for (int i = 0; i < 4; ++i)
    if (ls[i]->N != HW) return false;
  long long elems = 0;
  for (int i = 0; i < a.G.n; ++i) elems = a.G.g[i].base + (long long)a.G.g[i].B * a.G.g[i].D * a.G.g[i].H * a.G.g[i].W * D.C;
  return elems < (1ll << 31);
}
template <int C, bool SCAF>
static int lp_tcw_render_forward_t(cudaStream_t st, const LpRenderArgs& a, const float* params, float* out_len, float* out_nlt,
                                   float* out_feat, int feat_stride) {
  const int groups = 2;
  const size_t bytes = WImg<C>::END + 128;
  if (LP_TC_SET_SMEM((lp_render_fwd_tcw_kernel<C, SCAF>), bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + GT - 1) / GT;
  int blocks = (tiles + groups - 1) / groups;
  if (blocks > lp_tc_num_sms()) blocks = lp_tc_num_sms();
  LP_LAUNCH((lp_render_fwd_tcw_kernel<C, SCAF>), dim3(blocks), dim3(groups * GT), bytes, st, a.R, a.M, a.D, a.G, a.SC, params, out_len,
            out_nlt, out_feat, feat_stride);
  return LP_OK;
}
static inline int lp_tcw_render_forward(cudaStream_t st, const LpRenderArgs& a, const float* params, float* out_len,
                                        float* out_nlt, float* out_feat, int feat_stride) {
  if (a.use_scaffold)
    return a.D.C == 16 ? lp_tcw_render_forward_t<16, true>(st, a, params, out_len, out_nlt, out_feat, feat_stride)
                       : lp_tcw_render_forward_t<32, true>(st, a, params, out_len, out_nlt, out_feat, feat_stride);
  return a.D.C == 16 ? lp_tcw_render_forward_t<16, false>(st, a, params, out_len, out_nlt, out_feat, feat_stride)
                     : lp_tcw_render_forward_t<32, false>(st, a, params, out_len, out_nlt, out_feat, feat_stride);
}


// ===========================================================================================
// backward, hidden width 64
// ===========================================================================================
// One group of 256 threads per SM: two threads per sample (thread part h owns columns [32h, 32h+32) of every 64-wide
// row and half of the grid channels), so that the per-thread code and register budget are those of the width-32
// kernel.  What differs from lp_render_bwd_ws_kernel, all forced by one SM's shared memory (227 KB) and tensor
// memory (512 columns):
//  * no transposed weight copies: the input-gradient products read the FORWARD weight tiles as MN-major operands
//    (lp_tc_mma_ts_t); opacity and colour hidden layers are separate products (D is 64 columns);
//  * the parameter-gradient products are transposed: the gradient tiles are the M side (two stacks of 128 rows,
//    [d_ho | d_hc] and [d_t | d_h1]), the activation tiles the N side ([trunk | ones], [h1 | x0 | ones]), so the
//    accumulators hold dW^T in 80 + 112 columns; last layer [ho | hc]^T x dYL as before (its 4 bias gradients are
//    summed per thread); the encoding's share of dWc0 is [.. | S]^T x enc per ray tile (64 columns).  464 columns.
template <int C>
struct WBImg {
  using I = WImg<C>;
  static constexpr int BARS = I::END;
  static constexpr int XCH = BARS + 128;                 // float4 [2][128]
  static constexpr int TILES = XCH + 4096;
  // operand tiles: chunk index (2048 B = 8 features x 128 samples) of each block
  static constexpr int H1 = 0, X0 = 8, ONES_A = X0 + C / 8, TR = ONES_A + 2, ONES_B = TR + 8, HO = ONES_B + 2, HC = HO + 8,
                       DHO = HC + 8, DHC = DHO + 8, DT = DHC + 8, DH1 = DT + 8, DYL = DH1 + 8, NCH = DYL + 2;
  static constexpr int BYTES = TILES + NCH * 2048;
};
constexpr int WB_A = 0, WB_E = 64, WB_D = 128;                       // group columns: A hi 0..31 / lo 32..63, encoding, D
constexpr int WB_P1 = 192, WB_P2 = 272, WB_P3 = 384, WB_P4 = 400;   // accumulators: 80 + 112 + 16 + 64 columns

template <int C>
LP_DEVICE void lp_wb_issue_dw(unsigned tmem, unsigned char* tl, int accumulate, int wi) {
  using B = WBImg<C>;
  const lp_kdesc_t a1 = lp_tc_mndesc_lo(tl + B::DHO * 2048), b1 = lp_tc_mndesc_lo(tl + B::TR * 2048),
                   a2 = lp_tc_mndesc_lo(tl + B::DT * 2048), b2 = lp_tc_mndesc_lo(tl + B::H1 * 2048),
                   a3 = lp_tc_mndesc_lo(tl + B::HO * 2048), b3 = lp_tc_mndesc_lo(tl + B::DYL * 2048);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    if (wi < 0 || (ks >> 1) == wi) {
      lp_tc_mma_ss_mn(tmem + WB_P1, lp_tc_kadv(a1, ks * 256), lp_tc_kadv(b1, ks * 256), 2048, 80, accumulate | (ks > 0));
      lp_tc_mma_ss_mn(tmem + WB_P2, lp_tc_kadv(a2, ks * 256), lp_tc_kadv(b2, ks * 256), 2048, 64 + C + 16, accumulate | (ks > 0));
      lp_tc_mma_ss_mn(tmem + WB_P3, lp_tc_kadv(a3, ks * 256), lp_tc_kadv(b3, ks * 256), 2048, 16, accumulate | (ks > 0));
    }
  }
}
template <int C>
LP_DEVICE void lp_wb_issue_encw(unsigned tmem, unsigned char* tl, int accumulate, int wi) {
  using B = WBImg<C>;
  const lp_kdesc_t a1 = lp_tc_mndesc_lo(tl + B::DHO * 2048), b = lp_tc_mndesc_lo(tl + B::H1 * 2048);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    if (wi < 0 || (ks >> 1) == wi)
      lp_tc_mma_ss_mn(tmem + WB_P4, lp_tc_kadv(a1, ks * 256), lp_tc_kadv(b, ks * 256), 2048, 64, accumulate | (ks > 0));
}
// issuer wi: k-step wi (16 outputs = two n-chunks of the tile) of an input-gradient product on a transposed weight view
LP_DEVICE void lp_issue_dx_part(unsigned tbase, int d_col, int a_col, const unsigned char* whi, const unsigned char* wlo, int nstride,
                                int n, int lo_off, int wi) {
  if (wi >= 0 && wi < 4) {
    const lp_kdesc_t bh = lp_tc_kdesc_lo_t(whi + wi * 2 * nstride, nstride), bl = lp_tc_kdesc_lo_t(wlo + wi * 2 * nstride, nstride);
    lp_tc_mma_ts_t(tbase + d_col, tbase + a_col + 8 * wi, bh, nstride, n, 1);
    lp_tc_mma_ts_t(tbase + d_col, tbase + a_col + lo_off + 8 * wi, bh, nstride, n, 1);
    lp_tc_mma_ts_t(tbase + d_col, tbase + a_col + 8 * wi, bl, nstride, n, 1);
  }
}

template <int C, bool SCAF>
__global__ void __launch_bounds__(256, 1) lp_render_bwd_tcw_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G, LpGridSet SC,
                                                                    const float* __restrict__ params, LpBwdIo io) {
  using I = WImg<C>;
  using B = WBImg<C>;
  constexpr int W = 32, CW = C / 2, GTH = 256, NS = (HW / 8) * 128, NS0 = (C / 8) * 128;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = tid % GT, h = tid / GT, wig = warp & 3, wi = warp;  // sample row, column part, TMEM lane quarter, issuer id
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + B::BARS);  // [0] round trips, [1] dW, [2] start-up
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 4);
  unsigned char* tl = sm + B::TILES;
  float4* xch = reinterpret_cast<float4*>(sm + B::XCH);
  lp_build_wimg<C>(sm, params, D);
  for (int e = tid; e < B::NCH * 128; e += GTH) reinterpret_cast<uint4*>(tl)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (h == 0) {  // rows of ones (bf16 1.0): first row of the two ones blocks
    *reinterpret_cast<unsigned short*>(tl + B::ONES_A * 2048 + (s >> 3) * 128 + (s & 7) * 16) = 0x3F80;
    *reinterpret_cast<unsigned short*>(tl + B::ONES_B * 2048 + (s >> 3) * 128 + (s & 7) * 16) = 0x3F80;
  }
  if (tid == 0) {
    lp_mbar_init(bars + 0, 4);
    lp_mbar_init(bars + 1, 4);
    lp_mbar_init(bars + 2, 1);
    lp_mbar_init_fence();
  }
  if (tid < 32) lp_tmem_alloc512(tmem_slot);
  lp_fence_async_smem();
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  const unsigned tmem = *tmem_slot;
  if (tid == 0) {  // clear the accumulators (all gradient tiles are zero)
    lp_wb_issue_dw<C>(tmem, tl, 0, -1);
    lp_wb_issue_encw<C>(tmem, tl, 0, -1);
    lp_tc_commit(bars + 2);
  }
  lp_mbar_wait(bars + 2, 0);
  lp_tc_fence_after();
  __syncthreads();

  const unsigned tbase = tmem;
  const unsigned tme = lp_taddr(tbase, wig, 0);
  const bool issuer = lane == 0 && warp < 4;
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  const int pk = 16 * h, fc = 32 * h, ck = 4 * h;  // packed-column / fp32-column / tile-chunk offset of this part
  lp_tmem_zero<32>(tme + WB_D + fc);
  const lp_kdesc_t w_t0h = lp_tc_kdesc_lo(sm + I::T0_HI), w_t0l = lp_tc_kdesc_lo(sm + I::T0_LO),
                   w_t1h = lp_tc_kdesc_lo(sm + I::T1_HI), w_t1l = lp_tc_kdesc_lo(sm + I::T1_LO),
                   w_och = lp_tc_kdesc_lo(sm + I::OC_HI), w_ocl = lp_tc_kdesc_lo(sm + I::OC_LO);
  unsigned long long *bar = bars, *bar_dw = bars + 1;
  int phase = 0, n_dw = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;
  float bl0 = 0.f, bl1 = 0.f, bl2 = 0.f, bl3 = 0.f;  // last-layer bias gradients of this thread's samples (part 0 only)

#define LP_WB_HANDOFF(ISSUE) LP_TCG_HANDOFF(1, GTH, (warp < 4 && lp_elect_one()), ISSUE)
#define LP_WB_WAIT() LP_TCG_WAIT(bar, phase)
#define LP_WB_ROUND(ISSUE) LP_WB_HANDOFF(ISSUE) LP_WB_WAIT()
  // this part's 32 columns of the accumulator row
#define LP_WB_LD(v) lp_tmem_ld32u(tme + WB_D + fc, v); lp_tmem_zero<32>(tme + WB_D + fc)

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const Ray1 me = lp_load_ray1(R, tile * GT + s, G.g[0].B);
    const int q = me.active ? me.ray : R.n - 1;
    {
      float e[W];
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)q * HW + fc);
#pragma unroll
      for (int k = 0; k < W / 4; ++k) {
        const float4 v = __ldg(e4 + k);
        e[4 * k] = v.x; e[4 * k + 1] = v.y; e[4 * k + 2] = v.z; e[4 * k + 3] = v.w;
      }
      lp_stage_row<W, 32>(tme + WB_E + pk, e);
    }
    LpCompBwd cb;
    cb.init(io, q, me.active, D.n_feat);
    float S[W];
#pragma unroll
    for (int j = 0; j < W; ++j) S[j] = 0.f;

    struct Pos { float depth, delta, x, y, z, oob, occ; };
    auto sample_at = [&](int step) {
      Pos p;
      const Sched sc = lp_sched(step, M);
      lp_depth_delta(sc, me.near, me.far, p.depth, p.delta);
      p.x = me.ox + p.depth * me.dx; p.y = me.oy + p.depth * me.dy; p.z = me.oz + p.depth * me.dz;
      if (M.contract) lp_contract(p.x, p.y, p.z);
      p.oob = M.mask_oob ? lp_in_bounds(p.x, p.y, p.z) : 1.f;
      p.occ = SCAF ? lp_nearest(SC, me.b, p.x, p.y, p.z) : 1.f;
      return p;
    };
    Pos cur = sample_at(0), prev = cur;
    float x0[CW], dxp[CW];
    bool cur_hit = lp_gather_regs<C, CW>(G, me.b, cur.x, cur.y, cur.z, cur.oob, x0, CW * h);
    bool have_prev = false;
    float e_raw = 0.f, e_lg0 = 0.f, e_lg1 = 0.f, e_lg2 = 0.f, G_raw = 0.f, L0 = 0.f, L1 = 0.f, L2 = 0.f;  // empty-space folding
    bool any_empty = false;

    auto composite = [&](float raw, float lg0, float lg1, float lg2, int step, float& g_raw, float& dl0, float& dl1, float& dl2) {
      cb.grad(M, me.ray, step, step == tot - 1, raw, lg0, lg1, lg2, cur.depth, cur.delta, SCAF ? cur.occ : 1.f, g_raw, dl0, dl1, dl2);
    };

    for (int step = LP_TC_EMPTY_FOLD ? -1 : 0; step < tot + (LP_TC_EMPTY_FOLD ? 1 : 0); ++step) {
      const bool probe = step < 0, virt = step == tot, real = !probe && !virt;
      if (virt && !any_empty) break;
      float v[W];
      if (SCAF && real && !lp_bar_any(1, GTH, cur.occ != 0.f)) {
        if (step + 1 < tot) {
          cur = sample_at(step + 1);
          cur_hit = lp_gather_regs<C, CW>(G, me.b, cur.x, cur.y, cur.z, cur.oob, x0, CW * h);
        }
        continue;
      }
      if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);  // the previous step's dW products have consumed the tiles
      if (real) {
        lp_tile_row<CW>(tl, B::X0 + (CW / 8) * h, s, x0);
        lp_stage_row<CW, 32>(tme + WB_A + (CW / 2) * h, x0);
      } else {
        float z0[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) z0[c] = 0.f;
        lp_tile_row<CW>(tl, B::X0 + (CW / 8) * h, s, z0);
        lp_stage_row<CW, 32>(tme + WB_A + (CW / 2) * h, z0);
      }
      // ------------------------------ forward recompute ------------------------------
      lp_tmem_wait_st();
      lp_tc_fence_before();
      if (!((LP_TC_EMPTY_FOLD ? lp_bar_any(1, GTH, real && cur_hit) : (lp_bar_sync(1, GTH), true)) || !real)) {
        float g_raw, dl0, dl1, dl2;  // every sample of the group is empty
        composite(e_raw, e_lg0, e_lg1, e_lg2, step, g_raw, dl0, dl1, dl2);
        G_raw += g_raw; L0 += dl0; L1 += dl1; L2 += dl2;
        any_empty = true;
        if (step + 1 < tot) {
          cur = sample_at(step + 1);
          cur_hit = lp_gather_regs<C, CW>(G, me.b, cur.x, cur.y, cur.z, cur.oob, x0, CW * h);
        }
        continue;
      }
      if (issuer) {
        lp_tc_fence_after();
        lp_issue_layer_part(tbase, WB_D, WB_A, w_t0h, w_t0l, C / 16, 0, NS0, HW, 32, wi);
        lp_tc_commit(bar);
      }
      if (have_prev) {
        if (me.active && prev.oob != 0.f) lp_splat_regs<C, CW>(G, io.g_grid, me.b, prev.x, prev.y, prev.z, dxp, CW * h);
        have_prev = false;
      }
      LP_WB_WAIT();
      LP_WB_LD(v);
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = fmaxf(v[j] + F[I::FB + fc + j], 0.f);
      lp_tile_row<W>(tl, B::H1 + ck, s, v);
      lp_stage_row<W, 32>(tme + WB_A + pk, v);
      LP_WB_ROUND(lp_issue_layer_part(tbase, WB_D, WB_A, w_t1h, w_t1l, 4, 0, NS, HW, 32, wi); lp_tc_commit(bar));
      LP_WB_LD(v);
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = fmaxf(v[j] + F[I::FB + HW + fc + j], 0.f);
      lp_tile_row<W>(tl, B::TR + ck, s, v);
      lp_stage_row<W, 32>(tme + WB_A + pk, v);
      // opacity hidden layer
      LP_WB_ROUND(lp_issue_layer_part(tbase, WB_D, WB_A, w_och, w_ocl, 4, 0, NS, HW, 32, wi); lp_tc_commit(bar));
      float raw = 0.f, lg0 = 0.f, lg1 = 0.f, lg2 = 0.f;
      LP_WB_LD(v);
      // colour hidden layer on the same staged trunk row (+ the encoding): issued before the opacity epilogue runs
      LP_WB_HANDOFF(lp_issue_layer_part(tbase, WB_D, WB_A, lp_tc_kadv(w_och, (HW / 8) * NS), lp_tc_kadv(w_ocl, (HW / 8) * NS), 4, 0, NS,
                                        HW, 32, wi);
                    lp_issue_layer_part(tbase, WB_D, WB_E, lp_tc_kadv(w_och, (HW / 8) * NS), lp_tc_kadv(w_ocl, (HW / 8) * NS), 4, 0, NS,
                                        HW, 32, wi);
                    lp_tc_commit(bar));
#pragma unroll
      for (int j = 0; j < W; ++j) {
        v[j] = fmaxf(v[j] + F[I::FB + 2 * HW + fc + j], 0.f);
        raw = fmaf(v[j], F[I::FWO + fc + j], raw);
      }
      lp_tile_row<W>(tl, B::HO + ck, s, v);
      LP_WB_WAIT();
      LP_WB_LD(v);
#pragma unroll
      for (int j = 0; j < W; ++j) {
        v[j] = fmaxf(v[j] + F[I::FB + 3 * HW + fc + j], 0.f);
        const float4 w = *reinterpret_cast<const float4*>(F + I::FWC + 4 * (fc + j));
        lg0 = fmaf(v[j], w.x, lg0); lg1 = fmaf(v[j], w.y, lg1); lg2 = fmaf(v[j], w.z, lg2);
      }
      lp_tile_row<W>(tl, B::HC + ck, s, v);
      {  // combine the two parts' partial sums of the output layer
        xch[h * GT + s] = make_float4(lg0, lg1, lg2, raw);
        lp_bar_sync(1, GTH);
        const float4 o = xch[(h ^ 1) * GT + s];
        lg0 += o.x; lg1 += o.y; lg2 += o.z; raw += o.w;
      }
      raw += F[I::FBL + 3]; lg0 += F[I::FBL]; lg1 += F[I::FBL + 1]; lg2 += F[I::FBL + 2];
      if (probe) {
        e_raw = raw; e_lg0 = lg0; e_lg1 = lg1; e_lg2 = lg2;
        continue;
      }
      // ------------------------------ compositing gradient ------------------------------
      float g_raw, dl0, dl1, dl2;
      if (!virt) composite(raw, lg0, lg1, lg2, step, g_raw, dl0, dl1, dl2);
      else { g_raw = G_raw; dl0 = L0; dl1 = L1; dl2 = L2; }
      if (h == 0) {
        lp_tile8(tl, B::DYL, s, dl0, dl1, dl2, g_raw, 0.f, 0.f, 0.f, 0.f);
        bl0 += dl0; bl1 += dl1; bl2 += dl2; bl3 += g_raw;
      }
      // ------------------------------ backward sweep ------------------------------
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = g_raw * F[I::FWO + fc + j];  // d_ho
      lp_gate_row<W>(v, tl, B::HO + ck, s);
      lp_tile_row<W>(tl, B::DHO + ck, s, v);
      lp_stage_row<W, 32>(tme + WB_A + pk, v);
      LP_WB_HANDOFF(lp_issue_dx_part(tbase, WB_D, WB_A, sm + I::OC_HI, sm + I::OC_LO, NS, HW, 32, wi); lp_tc_commit(bar));
#pragma unroll
      for (int j = 0; j < W; ++j) {                                     // d_hc (while d_ho Wo0^T runs)
        const float4 w = *reinterpret_cast<const float4*>(F + I::FWC + 4 * (fc + j));
        v[j] = fmaf(dl0, w.x, fmaf(dl1, w.y, dl2 * w.z));
      }
      lp_gate_row<W>(v, tl, B::HC + ck, s);
#pragma unroll
      for (int j = 0; j < W; ++j) S[j] += v[j];
      lp_tile_row<W>(tl, B::DHC + ck, s, v);
      LP_WB_WAIT();  // the operand columns are free again
      lp_stage_row<W, 32>(tme + WB_A + pk, v);
      LP_WB_HANDOFF(lp_issue_dx_part(tbase, WB_D, WB_A, sm + I::OC_HI + (HW / 8) * NS, sm + I::OC_LO + (HW / 8) * NS, NS, HW, 32, wi);
                    lp_tc_commit(bar));
      if (!virt) prev = cur;
      if (step + 1 < tot) {  // prefetch the next step's features while the product runs
        cur = sample_at(step + 1);
        cur_hit = lp_gather_regs<C, CW>(G, me.b, cur.x, cur.y, cur.z, cur.oob, x0, CW * h);
      }
      LP_WB_WAIT();
      LP_WB_LD(v);  // d_t = d_ho Wo0^T + d_hc Wc0^T (accumulated in D)
      lp_gate_row<W>(v, tl, B::TR + ck, s);
      lp_tile_row<W>(tl, B::DT + ck, s, v);
      lp_stage_row<W, 32>(tme + WB_A + pk, v);
      LP_WB_ROUND(lp_issue_dx_part(tbase, WB_D, WB_A, sm + I::T1_HI, sm + I::T1_LO, NS, HW, 32, wi); lp_tc_commit(bar));
      LP_WB_LD(v);
      lp_gate_row<W>(v, tl, B::H1 + ck, s);  // d_h1
      lp_tile_row<W>(tl, B::DH1 + ck, s, v);
      lp_stage_row<W, 32>(tme + WB_A + pk, v);
      lp_fence_async_smem();  // this step's tile writes -> visible to the tensor core
      LP_WB_ROUND(lp_issue_dx_part(tbase, WB_D, WB_A, sm + I::T0_HI, sm + I::T0_LO, NS0, C, 32, wi); lp_tc_commit(bar);
                  lp_wb_issue_dw<C>(tmem, tl, 1, wi); lp_tc_commit(bar_dw));
      ++n_dw;
      lp_tmem_ld<CW>(tme + WB_D + CW * h, dxp);
      lp_tmem_zero<CW>(tme + WB_D + CW * h);
#pragma unroll
      for (int c = 0; c < CW; ++c) dxp[c] *= prev.oob;
      have_prev = !virt;
    }
    if (have_prev && me.active && prev.oob != 0.f)
      lp_splat_regs<C, CW>(G, io.g_grid, me.b, prev.x, prev.y, prev.z, dxp, CW * h);
    // ---- per-tile tail: encoding gradient = S Wc0^T, and the encoding's share of dWc0^T = S^T enc ----
    if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);
    {
      float e[W];
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)q * HW + fc);
#pragma unroll
      for (int k = 0; k < W / 4; ++k) {
        const float4 vv = __ldg(e4 + k);
        e[4 * k] = vv.x; e[4 * k + 1] = vv.y; e[4 * k + 2] = vv.z; e[4 * k + 3] = vv.w;
      }
      lp_tile_row<W>(tl, B::H1 + ck, s, e);
      lp_tile_row<W>(tl, B::DHC + ck, s, S);
      lp_stage_row<W, 32>(tme + WB_A + pk, S);
      lp_fence_async_smem();
      float v[W];
      LP_WB_ROUND(lp_issue_dx_part(tbase, WB_D, WB_A, sm + I::OC_HI + (HW / 8) * NS, sm + I::OC_LO + (HW / 8) * NS, NS, HW, 32, wi);
                  lp_tc_commit(bar); lp_wb_issue_encw<C>(tmem, tl, 1, wi); lp_tc_commit(bar_dw));
      ++n_dw;
      LP_WB_LD(v);
      if (me.active) {
        float4* ge = reinterpret_cast<float4*>(io.g_enc + (long long)me.ray * HW + fc);
#pragma unroll
        for (int k = 0; k < W / 4; ++k) ge[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      }
    }
  }
#undef LP_WB_ROUND
#undef LP_WB_HANDOFF
#undef LP_WB_WAIT
#undef LP_WB_LD
  if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  {
    const LpLayer &t0 = D.trunk.l[0], &t1 = D.trunk.l[1], &o0 = D.opacity.l[0], &o1 = D.opacity.l[1],
                  &c0 = D.color.l[0], &c1 = D.color.l[1];
    if (h == 0) {  // last-layer bias gradients, summed per thread
      for (int c = 0; c < D.n_feat; ++c) lp_red_add1(io.g_params + c1.b_off + c, c == 0 ? bl0 : (c == 1 ? bl1 : bl2));
      lp_red_add1(io.g_params + o1.b_off, bl3);
    }
    if (warp < 4) {  // TMEM lane = stack row (a GRADIENT feature j), columns = input features i: the accumulators hold dW^T
      float v[32];
      const unsigned tlane = lp_taddr(tmem, warp, 0);
      const int row = 32 * warp + lane, j = row & 63;
      const bool lo_half = row < 64;
      // P1: rows [d_ho | d_hc] x cols [trunk (64) | ones]
      const LpLayer& L1 = lo_half ? o0 : c0;
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        lp_tmem_ld32u(tlane + WB_P1 + 32 * blk, v);
        for (int i = 0; i < 32; ++i) lp_red_add1(io.g_params + L1.w_off + (32 * blk + i) * L1.N + j, v[i]);
      }
      lp_tmem_ld32u(tlane + WB_P1 + 64, v);  // 16 valid columns: [ones, 0...]
      lp_red_add1(io.g_params + L1.b_off + j, v[0]);
      if (!lo_half) {  // P4: the encoding's share of dWc0^T (rows d_hc slot = S)
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          lp_tmem_ld32u(tlane + WB_P4 + 32 * blk, v);
          for (int i = 0; i < 32; ++i) lp_red_add1(io.g_params + c0.w_off + (32 * blk + i) * c0.N + j, v[i]);
        }
      }
      // P2: rows [d_t | d_h1] x cols [h1 (64) | x0 (C) | ones]
      if (lo_half) {
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          lp_tmem_ld32u(tlane + WB_P2 + 32 * blk, v);
          for (int i = 0; i < 32; ++i) lp_red_add1(io.g_params + t1.w_off + (32 * blk + i) * t1.N + j, v[i]);
        }
      } else {
        lp_tmem_ld32u(tlane + WB_P2 + 64, v);
        for (int i = 0; i < C && i < 32; ++i) lp_red_add1(io.g_params + t0.w_off + i * t0.N + j, v[i]);
      }
      lp_tmem_ld32u(tlane + WB_P2 + 64 + C - 16 * (C / 32), v);  // window containing the ones column (index 64 + C)
      lp_red_add1(io.g_params + (lo_half ? t1.b_off : t0.b_off) + j, v[16 * (C / 32)]);
      // P3: rows [ho | hc] x cols [dlogit_0..2, g_raw]
      lp_tmem_ld32u(tlane + WB_P3, v);  // 16 valid columns
      if (lo_half) lp_red_add1(io.g_params + o1.w_off + j * o1.N, v[3]);
      else
        for (int c = 0; c < D.n_feat; ++c) lp_red_add1(io.g_params + c1.w_off + j * c1.N + c, v[c]);
    }
  }
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(tmem);
}

template <int C, bool SCAF>
static int lp_tcw_render_backward_t(cudaStream_t st, const LpRenderArgs& a, const float* params, const LpBwdIo& io) {
  const size_t bytes = WBImg<C>::BYTES;
  if (LP_TC_SET_SMEM((lp_render_bwd_tcw_kernel<C, SCAF>), bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + GT - 1) / GT;
  int blocks = tiles;
  if (blocks > lp_tc_num_sms()) blocks = lp_tc_num_sms();
  LP_LAUNCH((lp_render_bwd_tcw_kernel<C, SCAF>), dim3(blocks), dim3(256), bytes, st, a.R, a.M, a.D, a.G, a.SC, params, io);
  return LP_OK;
}
static inline int lp_tcw_render_backward(cudaStream_t st, const LpRenderArgs& a, const float* params, const LpBwdIo& io) {
  if (a.use_scaffold)
    return a.D.C == 16 ? lp_tcw_render_backward_t<16, true>(st, a, params, io) : lp_tcw_render_backward_t<32, true>(st, a, params, io);
  return a.D.C == 16 ? lp_tcw_render_backward_t<16, false>(st, a, params, io) : lp_tcw_render_backward_t<32, false>(st, a, params, io);
}

}  // namespace lptc
