// Tensor-core Renderer path for the separate-colour-grid ("ReLU field") decoder shape: no trunk,
// opacity and colour MLPs of 2 layers x hidden 32 fed by relu(sample(grid)) and relu(sample(color_grid))
// + ray_encoding (reference: renderer_fw.py:267-316, renderer_bw.py).  Same thread-per-sample tcgen05
// scheme as lp_render_tc.cuh (read its header first), but the two hidden layers are ONE block-diagonal
// product  [x_o | x_c] (K = 2C)  x  diag(Wo0, Wc0) (N = 64), and the input gradient one product
// [d_ho | d_hc] (K = 64) x diag(Wo0^T, Wc0^T) (N = 2C): one tensor-core round trip per step forward,
// two backward.  Other colour-grid configurations (scaffold, other layer counts) take the generic kernels.
#pragma once

#include "lp_render_tc.cuh"

namespace lptc {

template <int C>
struct CgImg {
  static constexpr int K2 = 2 * C;
  static constexpr int OC_HI = 0;                       // [64 out: opacity hidden | colour hidden][2C in: grid | colour grid]
  static constexpr int OC_LO = OC_HI + 64 * K2 * 2;
  static constexpr int F32 = OC_LO + 64 * K2 * 2;       // fp32: b_o0[32] b_c0[32] | wo1[32] | Wc1[32][4] | b_last[4]
  static constexpr int FB = 0, FWO = 64, FWC = 96, FBL = 224, NF = 228;
  static constexpr int FWD_END = (F32 + NF * 4 + 127) / 128 * 128;
  // backward
  static constexpr int X_HI = FWD_END;                  // [2C: d grid | d colour grid][64: opacity hidden | colour hidden]
  static constexpr int X_LO = X_HI + K2 * 64 * 2;
  static constexpr int BARS = X_LO + K2 * 64 * 2;
  static constexpr int GROUPS = BARS + 128;
  static constexpr int ONES1 = K2 / 8;                   // A1 = [x_o | x_c | ones]
  static constexpr int A1 = 0;
  static constexpr int A2 = A1 + (ONES1 + 1) * 2048;     // [opacity hidden | colour hidden | ones]
  static constexpr int DY = A2 + 9 * 2048;               // [d_ho | d_hc]
  static constexpr int DYL = DY + 8 * 2048;
  static constexpr int GROUP_BYTES = DYL + 6 * 2048;     // (+ slack so that A2's 16-chunk operand window stays inside)
  static_assert(A2 + 16 * 2048 <= GROUP_BYTES, "operand window leaves the group's region");
};
// tensor-memory columns per group: A hi 0..31 / lo 32..63, D 64..127; backward accumulators behind the groups
constexpr int CG_A = 0, CG_D = 64, CG_GROUP_COLS = 128;

template <int C>
LP_DEVICE void lp_build_cgimg(unsigned char* sm, const float* __restrict__ P, const LpDecoder& D, bool with_dx) {
  using I = CgImg<C>;
  const LpLayer &o0 = D.opacity.l[0], &o1 = D.opacity.l[1], &c0 = D.color.l[0], &c1 = D.color.l[1];
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int e = tid; e < 64 * I::K2; e += nth) {
    const int n = e & 63, k = e >> 6;
    float w = 0.f;
    if (n < 32 && k < C) w = P[o0.w_off + k * o0.N + n];
    if (n >= 32 && k >= C) w = P[c0.w_off + (k - C) * c0.N + (n - 32)];
    lp_put_w(sm, I::OC_HI, I::OC_LO, n, k, I::K2, w);
    if (with_dx) lp_put_w(sm, I::X_HI, I::X_LO, k, n, 64, w);  // transposed image: B[n' = input k][k' = hidden n]
  }
  float* F = reinterpret_cast<float*>(sm + I::F32);
  for (int e = tid; e < 32; e += nth) {
    F[I::FB + e] = P[o0.b_off + e];
    F[I::FB + 32 + e] = P[c0.b_off + e];
    F[I::FWO + e] = P[o1.w_off + e * o1.N];
    for (int c = 0; c < 4; ++c) F[I::FWC + 4 * e + c] = c < D.n_feat ? P[c1.w_off + e * c1.N + c] : 0.f;
  }
  if (tid < 4) F[I::FBL + tid] = tid == 3 ? P[o1.b_off] : (tid < D.n_feat ? P[c1.b_off + tid] : 0.f);
}

// this thread's sample: relu(sample(G)) -> xo, relu(sample(CG)) + enc -> xc; bit c of the masks = (feature c > 0)
template <int C>
LP_DEVICE void lp_cg_inputs(const LpGridSet& G, const LpGridSet& CG, int b, float x, float y, float z, float oob,
                            const float (&enc)[C], float (&xo)[C], float (&xc)[C], unsigned& m_o, unsigned& m_c) {
  lp_gather_regs<C>(G, b, x, y, z, oob, xo);
  lp_gather_regs<C>(CG, b, x, y, z, oob, xc);
  m_o = 0; m_c = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    xo[c] = fmaxf(xo[c], 0.f);
    xc[c] = fmaxf(xc[c], 0.f);
    m_o |= (xo[c] > 0.f ? 1u : 0u) << c;
    m_c |= (xc[c] > 0.f ? 1u : 0u) << c;
    xc[c] += enc[c];
  }
}

// ===========================================================================================
// forward
// ===========================================================================================
template <int C>
__global__ void __launch_bounds__(512, 1) lp_render_fwd_cg_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G, LpGridSet CG,
                                                                   const float* __restrict__ params,
                                                                   float* __restrict__ out_len, float* __restrict__ out_nlt,
                                                                   float* __restrict__ out_feat, int feat_stride) {
  using I = CgImg<C>;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x;
  const int grp = tid / GT, ngroups = blockDim.x / GT, wig = (tid >> 5) & 3;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + I::FWD_END);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  lp_build_cgimg<C>(sm, params, D, false);
  if (tid == 0) {
    for (int i = 0; i < ngroups; ++i) lp_mbar_init(bars + i, 4);
    lp_mbar_init_fence();
  }
  if (tid < 32) lp_tmem_alloc512(tmem_slot);
  lp_fence_async_smem();
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  const unsigned tbase = *tmem_slot + (unsigned)(grp * CG_GROUP_COLS);
  const unsigned tme = lp_taddr(tbase, wig, 0);
  const bool issuer = (tid & 31) == 0;
  lp_tmem_zero<32>(tme + CG_D);
  lp_tmem_zero<32>(tme + CG_D + 32);
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  const lp_kdesc_t w_h = lp_tc_kdesc_lo(sm + I::OC_HI), w_l = lp_tc_kdesc_lo(sm + I::OC_LO);
  unsigned long long* bar = bars + grp;
  int phase = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;

  for (int tile = blockIdx.x * ngroups + grp; tile < num_tiles; tile += gridDim.x * ngroups) {
    const Ray1 me = lp_load_ray1(R, tile * GT + (tid % GT), G.g[0].B);
    float enc[C];
    {
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)(me.active ? me.ray : R.n - 1) * C);
#pragma unroll
      for (int k = 0; k < C / 4; ++k) {
        const float4 v = __ldg(e4 + k);
        enc[4 * k] = v.x; enc[4 * k + 1] = v.y; enc[4 * k + 2] = v.z; enc[4 * k + 3] = v.w;
      }
    }
    LpCompFwd cf;
    for (int step = 0; step < tot; ++step) {
      const Sched sc = lp_sched(step, M);
      float depth, delta;
      lp_depth_delta(sc, me.near, me.far, depth, delta);
      {
        float x = me.ox + depth * me.dx, y = me.oy + depth * me.dy, z = me.oz + depth * me.dz;
        if (M.contract) lp_contract(x, y, z);
        const float oob = M.mask_oob ? lp_in_bounds(x, y, z) : 1.f;
        float xo[C], xc[C];
        unsigned m_o, m_c;
        lp_cg_inputs<C>(G, CG, me.b, x, y, z, oob, enc, xo, xc, m_o, m_c);
        lp_stage_row<C, 32>(tme + CG_A, xo);
        lp_stage_row<C, 32>(tme + CG_A + C / 2, xc);
      }
      lp_tmem_wait_st();
      lp_tc_fence_before();
      lp_bar_sync(1 + grp, GT);
      if (issuer) {
        lp_tc_fence_after();
        lp_issue_layer_part(tbase, CG_D, CG_A, w_h, w_l, I::K2 / 16, 0, (I::K2 / 8) * 128, 64, 32, wig);
        lp_tc_commit(bar);
      }
      lp_mbar_wait(bar, phase); phase ^= 1;
      lp_tc_fence_after();
      float v[32];
      float raw = F[I::FBL + 3], lg0 = F[I::FBL], lg1 = F[I::FBL + 1], lg2 = F[I::FBL + 2];
      lp_tmem_ld32u(tme + CG_D, v);
      lp_tmem_zero<32>(tme + CG_D);
#pragma unroll
      for (int j = 0; j < 32; ++j) raw = fmaf(fmaxf(v[j] + F[I::FB + j], 0.f), F[I::FWO + j], raw);
      lp_tmem_ld32u(tme + CG_D + 32, v);
      lp_tmem_zero<32>(tme + CG_D + 32);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float hc = fmaxf(v[j] + F[I::FB + 32 + j], 0.f);
        const float4 w = *reinterpret_cast<const float4*>(F + I::FWC + 4 * j);
        lg0 = fmaf(hc, w.x, lg0); lg1 = fmaf(hc, w.y, lg1); lg2 = fmaf(hc, w.z, lg2);
      }
      cf.add(M, me.ray, step, raw, lg0, lg1, lg2, depth, delta, 1.f);
    }
    if (me.active) {
      out_len[me.ray] = cf.len;
      out_nlt[me.ray] = cf.nlt;
      for (int c = 0; c < D.n_feat; ++c) out_feat[(long long)me.ray * feat_stride + c] = c == 0 ? cf.c0 : (c == 1 ? cf.c1 : cf.c2);
    }
  }
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(*tmem_slot);
}

// ===========================================================================================
// backward (one thread per sample, three groups per CTA)
// ===========================================================================================
constexpr int CGB_W = 3 * CG_GROUP_COLS, CGB_L = CGB_W + 64;  // CTA-wide accumulators: A1^T DY (N = 64), A2^T DYL (N = 16)

template <int C>
LP_DEVICE void lp_cg_issue_dw(unsigned tmem, unsigned char* gs, int accumulate, int wi) {
  using I = CgImg<C>;
  const lp_kdesc_t a1 = lp_tc_mndesc_lo(gs + I::A1), a2 = lp_tc_mndesc_lo(gs + I::A2), dy = lp_tc_mndesc_lo(gs + I::DY),
                   dyl = lp_tc_mndesc_lo(gs + I::DYL);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    if (wi < 0 || (ks >> 1) == wi) {  // wi < 0: one thread issues everything (start-up clear)
      lp_tc_mma_ss_mn(tmem + CGB_W, lp_tc_kadv(a1, ks * 256), lp_tc_kadv(dy, ks * 256), 2048, 64, accumulate | (ks > 0));
      lp_tc_mma_ss_mn(tmem + CGB_L, lp_tc_kadv(a2, ks * 256), lp_tc_kadv(dyl, ks * 256), 2048, 16, accumulate | (ks > 0));
    }
  }
}

template <int C>
__global__ void __launch_bounds__(384, 1) lp_render_bwd_cg_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G, LpGridSet CG,
                                                                   const float* __restrict__ params, LpBwdIo io) {
  using I = CgImg<C>;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / GT, ngroups = blockDim.x / GT, s = tid % GT, wig = warp & 3;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + I::BARS);  // [2g] round trips, [2g+1] dW; [8] init
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 10);
  unsigned char* gs = sm + I::GROUPS + grp * I::GROUP_BYTES;
  lp_build_cgimg<C>(sm, params, D, true);
  for (int e = s; e < I::GROUP_BYTES / 16; e += GT) reinterpret_cast<uint4*>(gs)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  *reinterpret_cast<unsigned short*>(gs + I::A1 + I::ONES1 * 2048 + (s >> 3) * 128 + (s & 7) * 16) = 0x3F80;
  *reinterpret_cast<unsigned short*>(gs + I::A2 + 8 * 2048 + (s >> 3) * 128 + (s & 7) * 16) = 0x3F80;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) lp_mbar_init(bars + i, 4);
    lp_mbar_init(bars + 8, 1);
    lp_mbar_init_fence();
  }
  if (tid < 32) lp_tmem_alloc512(tmem_slot);
  lp_fence_async_smem();
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  const unsigned tmem = *tmem_slot;
  if (tid == 0) {  // clear the CTA's accumulators
    lp_cg_issue_dw<C>(tmem, gs, 0, -1);
    lp_tc_commit(bars + 8);
  }
  lp_mbar_wait(bars + 8, 0);
  lp_tc_fence_after();
  __syncthreads();

  const unsigned tbase = tmem + (unsigned)(grp * CG_GROUP_COLS);
  const unsigned tme = lp_taddr(tbase, wig, 0);
  const bool issuer = lane == 0;
  lp_tmem_zero<32>(tme + CG_D);
  lp_tmem_zero<32>(tme + CG_D + 32);
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  const lp_kdesc_t w_h = lp_tc_kdesc_lo(sm + I::OC_HI), w_l = lp_tc_kdesc_lo(sm + I::OC_LO),
                   x_h = lp_tc_kdesc_lo(sm + I::X_HI), x_l = lp_tc_kdesc_lo(sm + I::X_LO);
  unsigned long long *bar = bars + 2 * grp, *bar_dw = bars + 2 * grp + 1;
  int phase = 0, n_dw = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;

#define LP_CG_ROUND(ISSUE) LP_TCG_HANDOFF(1 + grp, GT, lp_elect_one(), ISSUE) LP_TCG_WAIT(bar, phase)

  for (int tile = blockIdx.x * ngroups + grp; tile < num_tiles; tile += gridDim.x * ngroups) {
    const Ray1 me = lp_load_ray1(R, tile * GT + s, G.g[0].B);
    const int q = me.active ? me.ray : R.n - 1;
    float enc[C], genc[C];
    {
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)q * C);
#pragma unroll
      for (int k = 0; k < C / 4; ++k) {
        const float4 v = __ldg(e4 + k);
        enc[4 * k] = v.x; enc[4 * k + 1] = v.y; enc[4 * k + 2] = v.z; enc[4 * k + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) genc[c] = 0.f;
    }
    LpCompBwd cb;
    cb.init(io, q, me.active, D.n_feat);

    for (int step = 0; step < tot; ++step) {
      const Sched sc = lp_sched(step, M);
      float depth, delta;
      lp_depth_delta(sc, me.near, me.far, depth, delta);
      float px = me.ox + depth * me.dx, py = me.oy + depth * me.dy, pz = me.oz + depth * me.dz;
      if (M.contract) lp_contract(px, py, pz);
      const float oob = M.mask_oob ? lp_in_bounds(px, py, pz) : 1.f;
      unsigned m_o, m_c;
      float v[32];
      {
        float xo[C], xc[C];
        lp_cg_inputs<C>(G, CG, me.b, px, py, pz, oob, enc, xo, xc, m_o, m_c);
        if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);  // previous step's dW products have consumed the tiles
        lp_tile_row<C>(gs + I::A1, 0, s, xo);
        lp_tile_row<C>(gs + I::A1, C / 8, s, xc);
        lp_stage_row<C, 32>(tme + CG_A, xo);
        lp_stage_row<C, 32>(tme + CG_A + C / 2, xc);
      }
      // ------------------------------ forward recompute ------------------------------
      LP_CG_ROUND(lp_issue_layer_part(tbase, CG_D, CG_A, w_h, w_l, I::K2 / 16, 0, (I::K2 / 8) * 128, 64, 32, wig); lp_tc_commit(bar));
      float raw = F[I::FBL + 3], lg0 = F[I::FBL], lg1 = F[I::FBL + 1], lg2 = F[I::FBL + 2];
      lp_tmem_ld32u(tme + CG_D, v);
      lp_tmem_zero<32>(tme + CG_D);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaxf(v[j] + F[I::FB + j], 0.f);
        raw = fmaf(v[j], F[I::FWO + j], raw);
      }
      lp_tile_row<32>(gs + I::A2, 0, s, v);
      lp_tmem_ld32u(tme + CG_D + 32, v);
      lp_tmem_zero<32>(tme + CG_D + 32);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaxf(v[j] + F[I::FB + 32 + j], 0.f);
        const float4 w = *reinterpret_cast<const float4*>(F + I::FWC + 4 * j);
        lg0 = fmaf(v[j], w.x, lg0); lg1 = fmaf(v[j], w.y, lg1); lg2 = fmaf(v[j], w.z, lg2);
      }
      lp_tile_row<32>(gs + I::A2, 4, s, v);
      // ------------------------------ compositing gradient (as lp_render_bwd_ws_kernel) ------------------------------
      float g_raw, dl0, dl1, dl2;
      cb.grad(M, me.ray, step, step == tot - 1, raw, lg0, lg1, lg2, depth, delta, 1.f, g_raw, dl0, dl1, dl2);
      lp_tile8(gs + I::DYL, 0, s, dl0, dl1, dl2, g_raw, 0.f, 0.f, 0.f, 0.f);
      // ------------------------------ backward sweep ------------------------------
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = g_raw * F[I::FWO + j];  // d_ho
      lp_gate_row<32>(v, gs + I::A2, 0, s);
      lp_tile_row<32>(gs + I::DY, 0, s, v);
      lp_stage_row<32, 32>(tme + CG_A, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {                               // d_hc
        const float4 w = *reinterpret_cast<const float4*>(F + I::FWC + 4 * j);
        v[j] = fmaf(dl0, w.x, fmaf(dl1, w.y, dl2 * w.z));
      }
      lp_gate_row<32>(v, gs + I::A2, 4, s);
      lp_tile_row<32>(gs + I::DY, 4, s, v);
      lp_stage_row<32, 32>(tme + CG_A + 16, v);
      lp_fence_async_smem();
      LP_CG_ROUND(lp_issue_layer_part(tbase, CG_D, CG_A, x_h, x_l, 4, 0, 1024, I::K2, 32, wig); lp_tc_commit(bar);
                  lp_cg_issue_dw<C>(tmem, gs, 1, wig); lp_tc_commit(bar_dw));
      ++n_dw;
      {
        float d[C];
        lp_tmem_ld<C>(tme + CG_D, d);          // gradient of relu(sample(grid))
        lp_tmem_zero<C>(tme + CG_D);
#pragma unroll
        for (int c = 0; c < C; ++c) d[c] = ((m_o >> c) & 1u) ? d[c] * oob : 0.f;
        if (me.active && oob != 0.f) lp_splat_regs<C>(G, io.g_grid, me.b, px, py, pz, d);
        lp_tmem_ld<C>(tme + CG_D + C, d);      // gradient of relu(sample(color_grid)) + encoding
        lp_tmem_zero<C>(tme + CG_D + C);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          genc[c] += d[c];
          d[c] = ((m_c >> c) & 1u) ? d[c] * oob : 0.f;
        }
        if (me.active && oob != 0.f) lp_splat_regs<C>(CG, io.g_cgrid, me.b, px, py, pz, d);
      }
    }
    if (me.active) {
      float4* ge = reinterpret_cast<float4*>(io.g_enc + (long long)me.ray * C);
#pragma unroll
      for (int k = 0; k < C / 4; ++k) ge[k] = make_float4(genc[4 * k], genc[4 * k + 1], genc[4 * k + 2], genc[4 * k + 3]);
    }
  }
#undef LP_CG_ROUND
  if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  if (warp < 4) {  // TMEM lane = stack row: A1 rows [x_o (C) | x_c (C) | ones], A2 rows [ho | hc | ones]
    const LpLayer &o0 = D.opacity.l[0], &o1 = D.opacity.l[1], &c0 = D.color.l[0], &c1 = D.color.l[1];
    float v[32];
    const unsigned tl = lp_taddr(tmem, warp, 0);
    const int row = 32 * warp + lane;
    lp_tmem_ld32u(tl + CGB_W, v);       // x d_ho
    if (row < C)
      for (int n = 0; n < 32; ++n) lp_red_add1(io.g_params + o0.w_off + row * o0.N + n, v[n]);
    if (row == 2 * C)
      for (int n = 0; n < 32; ++n) lp_red_add1(io.g_params + o0.b_off + n, v[n]);
    lp_tmem_ld32u(tl + CGB_W + 32, v);  // x d_hc
    if (row >= C && row < 2 * C)
      for (int n = 0; n < 32; ++n) lp_red_add1(io.g_params + c0.w_off + (row - C) * c0.N + n, v[n]);
    if (row == 2 * C)
      for (int n = 0; n < 32; ++n) lp_red_add1(io.g_params + c0.b_off + n, v[n]);
    if (warp < 3) {
      lp_tmem_ld32u(tl + CGB_L, v);     // 16 valid columns: dlogit_0..2, g_raw
      if (warp == 0) {
        lp_red_add1(io.g_params + o1.w_off + lane * o1.N, v[3]);
      } else if (warp == 1) {
        for (int c = 0; c < D.n_feat; ++c) lp_red_add1(io.g_params + c1.w_off + lane * c1.N + c, v[c]);
      } else if (lane == 0) {
        for (int c = 0; c < D.n_feat; ++c) lp_red_add1(io.g_params + c1.b_off + c, v[c]);
        lp_red_add1(io.g_params + o1.b_off, v[3]);
      }
    }
  }
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(tmem);
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
static inline bool lp_cg_render_supported(const LpRenderArgs& a) {
  const LpDecoder& D = a.D;
  if (!D.use_color_grid || a.use_scaffold) return false;
  if (D.trunk.n_layers != 0 || D.opacity.n_layers != 2 || D.color.n_layers != 2) return false;
  if (D.C != 16 && D.C != 32) return false;
  if (D.n_feat > 3 || D.in_c != D.C || a.CG.C != D.C) return false;
  if (D.opacity.l[0].N != H || D.color.l[0].N != H) return false;
  const LpGridSet* gs[2] = {&a.G, &a.CG};
  for (int j = 0; j < 2; ++j) {
    long long elems = 0;
    for (int i = 0; i < gs[j]->n; ++i) {
      const LpGrid& g = gs[j]->g[i];
      elems = g.base + (long long)g.B * g.D * g.H * g.W * D.C;
    }
    if (elems >= (1ll << 31)) return false;
  }
  return true;
}

template <int C>
static int lp_cg_render_forward_t(cudaStream_t st, const LpRenderArgs& a, const float* params, float* out_len, float* out_nlt,
                                  float* out_feat, int feat_stride) {
  const int groups = 4;
  const size_t bytes = CgImg<C>::FWD_END + 128;
  if (LP_TC_SET_SMEM(lp_render_fwd_cg_kernel<C>, bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + GT - 1) / GT;
  int blocks = (tiles + groups - 1) / groups;
  if (blocks > lp_tc_num_sms()) blocks = lp_tc_num_sms();
  LP_LAUNCH(lp_render_fwd_cg_kernel<C>, dim3(blocks), dim3(groups * GT), bytes, st, a.R, a.M, a.D, a.G, a.CG, params, out_len,
            out_nlt, out_feat, feat_stride);
  return LP_OK;
}
static inline int lp_cg_render_forward(cudaStream_t st, const LpRenderArgs& a, const float* params, float* out_len,
                                       float* out_nlt, float* out_feat, int feat_stride) {
  return a.D.C == 16 ? lp_cg_render_forward_t<16>(st, a, params, out_len, out_nlt, out_feat, feat_stride)
                     : lp_cg_render_forward_t<32>(st, a, params, out_len, out_nlt, out_feat, feat_stride);
}
template <int C>
static int lp_cg_render_backward_t(cudaStream_t st, const LpRenderArgs& a, const float* params, const LpBwdIo& io) {
  const int groups = C == 16 ? 3 : 2;  // shared memory: 56 / 64 KB of operand tiles per group
  const size_t bytes = CgImg<C>::GROUPS + (size_t)groups * CgImg<C>::GROUP_BYTES;
  if (LP_TC_SET_SMEM(lp_render_bwd_cg_kernel<C>, bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + GT - 1) / GT;
  int blocks = (tiles + groups - 1) / groups;
  if (blocks > lp_tc_num_sms()) blocks = lp_tc_num_sms();
  LP_LAUNCH(lp_render_bwd_cg_kernel<C>, dim3(blocks), dim3(groups * GT), bytes, st, a.R, a.M, a.D, a.G, a.CG, params, io);
  return LP_OK;
}
static inline int lp_cg_render_backward(cudaStream_t st, const LpRenderArgs& a, const float* params, const LpBwdIo& io) {
  return a.D.C == 16 ? lp_cg_render_backward_t<16>(st, a, params, io) : lp_cg_render_backward_t<32>(st, a, params, io);
}

}  // namespace lptc
