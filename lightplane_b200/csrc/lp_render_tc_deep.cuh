// Tensor-core Renderer kernels for decoders with layer counts other than 2/2/2 (hidden width 32, no separate colour
// grid): trunk of 1..4 layers, opacity and colour MLPs of 1..4 layers, at most 8 hidden (32-wide) layers in total.
// Same thread-per-sample tcgen05 scheme as lp_render_tc.cuh (read its header first), but table driven: the hidden
// layers are numbered  trunk 0..Lt-1 | opacity hidden | colour hidden  and walked by run-time loops, one tensor-core
// round trip per layer and direction.  The resource-saving forms of lp_render_tc_wide.cuh are used throughout: the
// input-gradient products read the forward weight tiles as MN-major operands, and the parameter-gradient products
// are transposed -- for every hidden layer l an M = 128 window starting at its gradient tile (rows 0..31 valid) times
// its input tile; bias gradients from windows of four gradient tiles times a tile of ones.  One group of 128 threads
// per SM in the backward (180 KB of operand tiles), two in the forward.
#pragma once

#include "lp_render_tc_wide.cuh"

namespace lptc {

constexpr int DP_MAXL = 8;
struct DeepPlan {
  int NL, Lt, Ho, Hc;                 // hidden layers: Lt trunk, Ho = Lo-1 opacity, Hc = Lc-1 colour
  int w_off[DP_MAXL], b_off[DP_MAXL], K[DP_MAXL];
  int wo_off, wo_N, bo_off, wc_off, wc_N, bc_off;  // output layers: opacity [32 -> 1], colour [32 -> n_feat]
  int n_feat;
};
LP_DEVICE int dp_o0(const DeepPlan& P) { return P.Lt; }                       // first / last hidden layer of a head
LP_DEVICE int dp_c0(const DeepPlan& P) { return P.Lt + P.Ho; }
LP_DEVICE int dp_src(const DeepPlan& P, int l) { return (l == dp_o0(P) || l == dp_c0(P)) ? P.Lt - 1 : l - 1; }  // -1: x0

// shared memory (bytes): weight tiles (hi at l*4096, lo at +2048; [32 out][K in] K-major) | fp32 section | barriers | tiles
struct DpImg {
  static constexpr int F32 = DP_MAXL * 4096;   // biases [8][32] | wo[32] | Wc[32][4] | b_last[4]
  static constexpr int FB = 0, FWO = 256, FWC = 288, FBL = 416, NF = 420;
  static constexpr int BARS = F32 + 1792;
  static constexpr int FWD_END = BARS;
  static constexpr int TILES = BARS + 128;
  // chunk (2048 B) indices of the backward's operand tiles
  static constexpr int X0 = 0, ACT = 4, LIN = 36, ONES = 44, DY = 46, DYL = 78, NCH = 90;
  static constexpr int BYTES = TILES + NCH * 2048;
};
// tensor-memory columns of a group: A hi 0..15 / lo 16..31, encoding hi 32..47 / lo 48..63, D 64..95
constexpr int DT_A = 0, DT_E = 32, DT_D = 64, DT_GROUP_COLS = 96;
constexpr int DT_ACC = 96, DT_ACCB = 352, DT_PL = 384, DT_PE = 400;  // backward accumulators (one group per SM)

LP_DEVICE void lp_build_dpimg(unsigned char* sm, const float* __restrict__ P, const DeepPlan& pl) {
  using I = DpImg;
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int l = 0; l < pl.NL; ++l) {
    const int K = pl.K[l];
    for (int e = tid; e < 32 * K; e += nth) {
      const int n = e & 31, k = e >> 5;
      lp_put_w(sm, l * 4096, l * 4096 + 2048, n, k, K, P[pl.w_off[l] + k * 32 + n]);
    }
  }
  float* F = reinterpret_cast<float*>(sm + I::F32);
  for (int e = tid; e < 32; e += nth) {
    for (int l = 0; l < pl.NL; ++l) F[I::FB + 32 * l + e] = P[pl.b_off[l] + e];
    F[I::FWO + e] = P[pl.wo_off + e * pl.wo_N];
    for (int c = 0; c < 4; ++c) F[I::FWC + 4 * e + c] = c < pl.n_feat ? P[pl.wc_off + e * pl.wc_N + c] : 0.f;
  }
  if (tid < 4) F[I::FBL + tid] = tid == 3 ? P[pl.bo_off] : (tid < pl.n_feat ? P[pl.bc_off + tid] : 0.f);
}

// issuer wi (lane 0 of warp wi of the group): k-step wi of a forward product on layer l's tile
LP_DEVICE void dp_issue_fwd(unsigned tbase, unsigned char* sm, int l, int K, int a_col, int wi) {
  lp_issue_layer_part(tbase, DT_D, a_col, lp_tc_kdesc_lo(sm + l * 4096), lp_tc_kdesc_lo(sm + l * 4096 + 2048), K / 16, 0, (K / 8) * 128, 32,
                      16, wi);
}
// ... of an input-gradient product (layer l's tile read transposed: K' = 32 outputs = 2 k-steps, N' = K inputs)
LP_DEVICE void dp_issue_dx(unsigned tbase, unsigned char* sm, int l, int K, int wi) {
  if (wi >= 0 && wi < 2) {
    const int ns = (K / 8) * 128;
    const lp_kdesc_t bh = lp_tc_kdesc_lo_t(sm + l * 4096 + wi * 2 * ns, ns), bl = lp_tc_kdesc_lo_t(sm + l * 4096 + 2048 + wi * 2 * ns, ns);
    lp_tc_mma_ts_t(tbase + DT_D, tbase + DT_A + 8 * wi, bh, ns, K, 1);
    lp_tc_mma_ts_t(tbase + DT_D, tbase + DT_A + 16 + 8 * wi, bh, ns, K, 1);
    lp_tc_mma_ts_t(tbase + DT_D, tbase + DT_A + 8 * wi, bl, ns, K, 1);
  }
}

// output layers on the CUDA cores
LP_DEVICE float dp_raw(const float* F, const float (&v)[32]) {
  float r = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) r = fmaf(v[j], F[DpImg::FWO + j], r);
  return r;
}
LP_DEVICE void dp_logits(const float* F, const float (&v)[32], float& l0, float& l1, float& l2) {
  l0 = l1 = l2 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float4 w = *reinterpret_cast<const float4*>(F + DpImg::FWC + 4 * j);
    l0 = fmaf(v[j], w.x, l0); l1 = fmaf(v[j], w.y, l1); l2 = fmaf(v[j], w.z, l2);
  }
}
LP_DEVICE void dp_load_enc(const LpRays& R, int q, float (&e)[32]) {
  const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)q * 32);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 v = __ldg(e4 + k);
    e[4 * k] = v.x; e[4 * k + 1] = v.y; e[4 * k + 2] = v.z; e[4 * k + 3] = v.w;
  }
}

#define LP_DP_ROUND(ISSUE) LP_TCG_HANDOFF(1 + grp, GT, lp_elect_one(), ISSUE; lp_tc_commit(bar)) LP_TCG_WAIT(bar, phase)
#define LP_DP_LD(v) lp_tmem_ld32u(tme + DT_D, v); lp_tmem_zero<32>(tme + DT_D)

// ===========================================================================================
// forward
// ===========================================================================================
template <int C, bool SCAF>
__global__ void __launch_bounds__(256, 1) lp_render_fwd_deep_kernel(LpRays R, LpMarch M, DeepPlan pl, LpGridSet G, LpGridSet SC,
                                                                     const float* __restrict__ params,
                                                                     float* __restrict__ out_len, float* __restrict__ out_nlt,
                                                                     float* __restrict__ out_feat, int feat_stride) {
  using I = DpImg;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x;
  const int grp = tid / GT, ngroups = blockDim.x / GT, wig = (tid >> 5) & 3, wi = wig;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + I::BARS);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  lp_build_dpimg(sm, params, pl);
  if (tid == 0) {
    for (int i = 0; i < ngroups; ++i) lp_mbar_init(bars + i, 4);
    lp_mbar_init_fence();
  }
  if (tid < 32) lp_tmem_alloc512(tmem_slot);
  lp_fence_async_smem();
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  const unsigned tbase = *tmem_slot + (unsigned)(grp * DT_GROUP_COLS);
  const unsigned tme = lp_taddr(tbase, wig, 0);
  const bool issuer = (tid & 31) == 0;
  lp_tmem_zero<32>(tme + DT_D);
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  unsigned long long* bar = bars + grp;
  int phase = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;
  const int o0 = dp_o0(pl), c0 = dp_c0(pl);

  for (int tile = blockIdx.x * ngroups + grp; tile < num_tiles; tile += gridDim.x * ngroups) {
    const Ray1 me = lp_load_ray1(R, tile * GT + (tid % GT), G.g[0].B);
    const int q = me.active ? me.ray : R.n - 1;
    {
      float e[32];
      dp_load_enc(R, q, e);
      lp_stage_row<32, 16>(tme + DT_E, e);
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
        if (LP_TC_EMPTY_FOLD && !probe && !lp_bar_any(1 + grp, GT, hit)) {  // every sample of the group is empty
          cf.add(M, me.ray, step, e_raw, e_lg0, e_lg1, e_lg2, depth, delta, occ);
          continue;
        }
        lp_stage_row<C, 16>(tme + DT_A, x0);
      }
      float v[32], trv[32];
      float raw = 0.f, lg0 = 0.f, lg1 = 0.f, lg2 = 0.f;
#pragma unroll 1
      for (int l = 0; l < pl.NL; ++l) {
        if (l == o0 || l == c0) lp_stage_row<32, 16>(tme + DT_A, trv);  // a head starts from the trunk output
        LP_DP_ROUND(dp_issue_fwd(tbase, sm, l, pl.K[l], DT_A, wi); if (l == c0) dp_issue_fwd(tbase, sm, l, 32, DT_E, wi));
        LP_DP_LD(v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + F[I::FB + 32 * l + j], 0.f);
        if (l == pl.Lt - 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) trv[j] = v[j];
        }
        if (l == o0 + pl.Ho - 1 && pl.Ho > 0) raw = dp_raw(F, v);
        if (l == pl.NL - 1 && pl.Hc > 0) dp_logits(F, v, lg0, lg1, lg2);
        lp_stage_row<32, 16>(tme + DT_A, v);
      }
      if (pl.Ho == 0) raw = dp_raw(F, trv);
      if (pl.Hc == 0) {
        float e[32];
        dp_load_enc(R, q, e);
#pragma unroll
        for (int j = 0; j < 32; ++j) e[j] += trv[j];
        dp_logits(F, e, lg0, lg1, lg2);
      }
      raw += F[I::FBL + 3]; lg0 += F[I::FBL]; lg1 += F[I::FBL + 1]; lg2 += F[I::FBL + 2];
      if (probe) { e_raw = raw; e_lg0 = lg0; e_lg1 = lg1; e_lg2 = lg2; continue; }
      cf.add(M, me.ray, step, raw, lg0, lg1, lg2, depth, delta, occ);
    }
    if (me.active) {
      out_len[me.ray] = cf.len;
      out_nlt[me.ray] = cf.nlt;
      for (int c = 0; c < pl.n_feat; ++c) out_feat[(long long)me.ray * feat_stride + c] = c == 0 ? cf.c0 : (c == 1 ? cf.c1 : cf.c2);
    }
  }
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(*tmem_slot);
}

// ===========================================================================================
// backward (one group of 128 threads per SM)
// ===========================================================================================
LP_DEVICE void dp_issue_dw(unsigned tmem, unsigned char* tl, const DeepPlan& pl, int C, int accumulate, int wi, bool enc_only) {
  using I = DpImg;
  // MMA number m of the step is issued by issuer m % 4 (wi < 0: by the calling thread)
  int m = 0;
  auto mine = [&]() { return wi < 0 || (m++ & 3) == wi; };
  if (!enc_only) {
    for (int l = 0; l < pl.NL; ++l) {
      const int src = dp_src(pl, l);
      const lp_kdesc_t a = lp_tc_mndesc_lo(tl + (I::DY + 4 * l) * 2048),
                       b = lp_tc_mndesc_lo(tl + (src < 0 ? I::X0 : I::ACT + 4 * src) * 2048);
      for (int ks = 0; ks < 8; ++ks)
        if (mine()) lp_tc_mma_ss_mn(tmem + DT_ACC + 32 * l, lp_tc_kadv(a, ks * 256), lp_tc_kadv(b, ks * 256), 2048, src < 0 ? C : 32, accumulate | (ks > 0));
    }
    for (int g = 0; g < 2; ++g) {  // bias gradients: four gradient tiles at a time x ones
      const lp_kdesc_t a = lp_tc_mndesc_lo(tl + (I::DY + 16 * g) * 2048), b = lp_tc_mndesc_lo(tl + I::ONES * 2048);
      for (int ks = 0; ks < 8; ++ks)
        if (mine()) lp_tc_mma_ss_mn(tmem + DT_ACCB + 16 * g, lp_tc_kadv(a, ks * 256), lp_tc_kadv(b, ks * 256), 2048, 16, accumulate | (ks > 0));
    }
    {  // output layers: [opacity input | colour input]^T x dYL
      const lp_kdesc_t a = lp_tc_mndesc_lo(tl + I::LIN * 2048), b = lp_tc_mndesc_lo(tl + I::DYL * 2048);
      for (int ks = 0; ks < 8; ++ks)
        if (mine()) lp_tc_mma_ss_mn(tmem + DT_PL, lp_tc_kadv(a, ks * 256), lp_tc_kadv(b, ks * 256), 2048, 16, accumulate | (ks > 0));
    }
  }
  if (enc_only || wi < 0) {  // per ray tile: (sum of the first colour layer's gradient)^T x encoding (held in ACT slot 0)
    const lp_kdesc_t a = lp_tc_mndesc_lo(tl + (I::DY + 4 * dp_c0(pl)) * 2048), b = lp_tc_mndesc_lo(tl + I::ACT * 2048);
    for (int ks = 0; ks < 8; ++ks)
      if (mine()) lp_tc_mma_ss_mn(tmem + DT_PE, lp_tc_kadv(a, ks * 256), lp_tc_kadv(b, ks * 256), 2048, 32, accumulate | (ks > 0));
  }
}

template <int C, bool SCAF>
__global__ void __launch_bounds__(128, 1) lp_render_bwd_deep_kernel(LpRays R, LpMarch M, DeepPlan pl, LpGridSet G, LpGridSet SC,
                                                                     const float* __restrict__ params, LpBwdIo io) {
  using I = DpImg;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = 0, s = tid, wig = warp, wi = warp;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + I::BARS);  // [0] round trips, [1] dW, [2] start-up
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 4);
  unsigned char* tl = sm + I::TILES;
  lp_build_dpimg(sm, params, pl);
  for (int e = tid; e < I::NCH * 128; e += GT) reinterpret_cast<uint4*>(tl)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  *reinterpret_cast<unsigned short*>(tl + I::ONES * 2048 + (s >> 3) * 128 + (s & 7) * 16) = 0x3F80;  // the row of ones
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
    dp_issue_dw(tmem, tl, pl, C, 0, -1, false);
    lp_tc_commit(bars + 2);
  }
  lp_mbar_wait(bars + 2, 0);
  lp_tc_fence_after();
  __syncthreads();

  const unsigned tbase = tmem;
  const unsigned tme = lp_taddr(tbase, wig, 0);
  const bool issuer = lane == 0;
  lp_tmem_zero<32>(tme + DT_D);
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  unsigned long long *bar = bars, *bar_dw = bars + 1;
  int phase = 0, n_dw = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;
  const int o0 = dp_o0(pl), c0 = dp_c0(pl);
  float bl0 = 0.f, bl1 = 0.f, bl2 = 0.f, bl3 = 0.f;  // output-layer bias gradients of this thread's samples

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const Ray1 me = lp_load_ray1(R, tile * GT + s, G.g[0].B);
    const int q = me.active ? me.ray : R.n - 1;
    {
      float e[32];
      dp_load_enc(R, q, e);
      lp_stage_row<32, 16>(tme + DT_E, e);
    }
    LpCompBwd cb;
    cb.init(io, q, me.active, pl.n_feat);
    float S[32];  // step-sum of the gradient at the colour branch's input side (see the tile tail)
#pragma unroll
    for (int j = 0; j < 32; ++j) S[j] = 0.f;

    // empty-space folding (see lp_render_bwd_ws_kernel): probe iteration, summed gradients of the empty steps, one fold iteration
    float e_raw = 0.f, e_lg0 = 0.f, e_lg1 = 0.f, e_lg2 = 0.f, G_raw = 0.f, L0 = 0.f, L1 = 0.f, L2 = 0.f;
    bool any_empty = false;
    for (int step = LP_TC_EMPTY_FOLD ? -1 : 0; step < tot + (LP_TC_EMPTY_FOLD ? 1 : 0); ++step) {
      const bool probe = step < 0, virt = step == tot, real = !probe && !virt;
      if (virt && !any_empty) break;
      const Sched sc = lp_sched(real ? step : 0, M);
      float depth, delta;
      lp_depth_delta(sc, me.near, me.far, depth, delta);
      float px = me.ox + depth * me.dx, py = me.oy + depth * me.dy, pz = me.oz + depth * me.dz;
      if (M.contract) lp_contract(px, py, pz);
      float occ = 1.f;
      if (SCAF && real) {
        occ = lp_nearest(SC, me.b, px, py, pz);
        if (!lp_bar_any(1, GT, occ != 0.f)) continue;
      }
      const float oob = M.mask_oob ? lp_in_bounds(px, py, pz) : 1.f;
      float v[32], trv[32];
      float raw = 0.f, lg0 = 0.f, lg1 = 0.f, lg2 = 0.f;
      {
        float x0[C];
        bool hit = false;
        if (real) {
          hit = lp_gather_regs<C>(G, me.b, px, py, pz, oob, x0);
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) x0[c] = 0.f;
        }
        if (LP_TC_EMPTY_FOLD && real && !lp_bar_any(1, GT, hit)) {  // every sample of the group is empty
          float g_raw, dl0, dl1, dl2;
          cb.grad(M, me.ray, step, step == tot - 1, e_raw, e_lg0, e_lg1, e_lg2, depth, delta, occ, g_raw, dl0, dl1, dl2);
          G_raw += g_raw; L0 += dl0; L1 += dl1; L2 += dl2;
          any_empty = true;
          continue;
        }
        if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);  // the previous step's dW products have consumed the tiles
        lp_tile_row<C>(tl, I::X0, s, x0);
        lp_stage_row<C, 16>(tme + DT_A, x0);
      }
      // ------------------------------ forward recompute ------------------------------
#pragma unroll 1
      for (int l = 0; l < pl.NL; ++l) {
        if (l == o0 || l == c0) lp_stage_row<32, 16>(tme + DT_A, trv);
        LP_DP_ROUND(dp_issue_fwd(tbase, sm, l, pl.K[l], DT_A, wi); if (l == c0) dp_issue_fwd(tbase, sm, l, 32, DT_E, wi));
        LP_DP_LD(v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + F[I::FB + 32 * l + j], 0.f);
        lp_tile_row<32>(tl, I::ACT + 4 * l, s, v);
        if (l == pl.Lt - 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) trv[j] = v[j];
        }
        if (l == o0 + pl.Ho - 1 && pl.Ho > 0) { raw = dp_raw(F, v); lp_tile_row<32>(tl, I::LIN, s, v); }
        if (l == pl.NL - 1 && pl.Hc > 0) { dp_logits(F, v, lg0, lg1, lg2); lp_tile_row<32>(tl, I::LIN + 4, s, v); }
        lp_stage_row<32, 16>(tme + DT_A, v);
      }
      if (pl.Ho == 0) { raw = dp_raw(F, trv); lp_tile_row<32>(tl, I::LIN, s, trv); }
      if (pl.Hc == 0) {
        float e[32];
        dp_load_enc(R, q, e);
#pragma unroll
        for (int j = 0; j < 32; ++j) e[j] += trv[j];
        dp_logits(F, e, lg0, lg1, lg2);
        lp_tile_row<32>(tl, I::LIN + 4, s, e);
      }
      raw += F[I::FBL + 3]; lg0 += F[I::FBL]; lg1 += F[I::FBL + 1]; lg2 += F[I::FBL + 2];
      // ------------------------------ compositing gradient (as lp_render_bwd_ws_kernel) ------------------------------
      if (probe) { e_raw = raw; e_lg0 = lg0; e_lg1 = lg1; e_lg2 = lg2; continue; }
      float g_raw, dl0, dl1, dl2;
      if (!virt) cb.grad(M, me.ray, step, step == tot - 1, raw, lg0, lg1, lg2, depth, delta, occ, g_raw, dl0, dl1, dl2);
      else { g_raw = G_raw; dl0 = L0; dl1 = L1; dl2 = L2; }
      lp_tile8(tl, I::DYL, s, dl0, dl1, dl2, g_raw, 0.f, 0.f, 0.f, 0.f);
      bl0 += dl0; bl1 += dl1; bl2 += dl2; bl3 += g_raw;
      // ------------------------------ backward sweep ------------------------------
      // opacity head: down to the gradient of its first hidden layer (kept in `dto`), or its direct term on the trunk
      float dto[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) dto[j] = g_raw * F[I::FWO + j];
      if (pl.Ho > 0) {
        lp_gate_row<32>(dto, tl, I::ACT + 4 * (o0 + pl.Ho - 1), s);
#pragma unroll 1
        for (int l = o0 + pl.Ho - 1; l > o0; --l) {
          lp_tile_row<32>(tl, I::DY + 4 * l, s, dto);
          lp_stage_row<32, 16>(tme + DT_A, dto);
          LP_DP_ROUND(dp_issue_dx(tbase, sm, l, 32, wi));
          LP_DP_LD(dto);
          lp_gate_row<32>(dto, tl, I::ACT + 4 * (l - 1), s);
        }
        lp_tile_row<32>(tl, I::DY + 4 * o0, s, dto);
      }
      // colour head
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(F + I::FWC + 4 * j);
        v[j] = fmaf(dl0, w.x, fmaf(dl1, w.y, dl2 * w.z));
      }
      if (pl.Hc > 0) {
        lp_gate_row<32>(v, tl, I::ACT + 4 * (pl.NL - 1), s);
#pragma unroll 1
        for (int l = pl.NL - 1; l > c0; --l) {
          lp_tile_row<32>(tl, I::DY + 4 * l, s, v);
          lp_stage_row<32, 16>(tme + DT_A, v);
          LP_DP_ROUND(dp_issue_dx(tbase, sm, l, 32, wi));
          LP_DP_LD(v);
          lp_gate_row<32>(v, tl, I::ACT + 4 * (l - 1), s);
        }
        lp_tile_row<32>(tl, I::DY + 4 * c0, s, v);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) S[j] += v[j];  // Hc > 0: gradient of the first colour layer; else of (trunk + encoding)
      // gradient of the trunk output: both heads' first layers transposed (accumulated in D) + the direct terms
      if (pl.Hc > 0) {
        lp_stage_row<32, 16>(tme + DT_A, v);
        LP_DP_ROUND(dp_issue_dx(tbase, sm, c0, 32, wi));
      }
      if (pl.Ho > 0) {
        lp_stage_row<32, 16>(tme + DT_A, dto);
        LP_DP_ROUND(dp_issue_dx(tbase, sm, o0, 32, wi));
      }
      {
        float d[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) d[j] = 0.f;
        if (pl.Hc > 0 || pl.Ho > 0) { LP_DP_LD(d); }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = d[j] + (pl.Hc == 0 ? v[j] : 0.f) + (pl.Ho == 0 ? dto[j] : 0.f);
      }
      lp_gate_row<32>(v, tl, I::ACT + 4 * (pl.Lt - 1), s);
      // trunk
#pragma unroll 1
      for (int l = pl.Lt - 1; l > 0; --l) {
        lp_tile_row<32>(tl, I::DY + 4 * l, s, v);
        lp_stage_row<32, 16>(tme + DT_A, v);
        LP_DP_ROUND(dp_issue_dx(tbase, sm, l, 32, wi));
        LP_DP_LD(v);
        lp_gate_row<32>(v, tl, I::ACT + 4 * (l - 1), s);
      }
      lp_tile_row<32>(tl, I::DY, s, v);
      lp_stage_row<32, 16>(tme + DT_A, v);
      lp_fence_async_smem();  // this step's tile writes -> visible to the tensor core
      lp_tmem_wait_st();
      lp_tc_fence_before();
      lp_bar_sync(1, GT);
      if (issuer) {
        lp_tc_fence_after();
        dp_issue_dx(tbase, sm, 0, C, wi);
        lp_tc_commit(bar);
        dp_issue_dw(tmem, tl, pl, C, 1, wi, false);
        lp_tc_commit(bar_dw);
      }
      lp_mbar_wait(bar, phase);
      phase ^= 1;
      lp_tc_fence_after();
      ++n_dw;
      {
        float d[C];
        lp_tmem_ld<C>(tme + DT_D, d);
        lp_tmem_zero<C>(tme + DT_D);
#pragma unroll
        for (int c = 0; c < C; ++c) d[c] *= oob;
        if (real && me.active && oob != 0.f) lp_splat_regs<C>(G, io.g_grid, me.b, px, py, pz, d);
      }
    }
    // ---- per-tile tail: encoding gradient and the encoding's share of the first colour layer's dW ----
    if (pl.Hc > 0) {
      if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);
      float e[32], v[32];
      dp_load_enc(R, q, e);
      lp_tile_row<32>(tl, I::ACT, s, e);
      lp_tile_row<32>(tl, I::DY + 4 * c0, s, S);
      lp_stage_row<32, 16>(tme + DT_A, S);
      lp_fence_async_smem();
      lp_tmem_wait_st();
      lp_tc_fence_before();
      lp_bar_sync(1, GT);
      if (issuer) {
        lp_tc_fence_after();
        dp_issue_dx(tbase, sm, c0, 32, wi);
        lp_tc_commit(bar);
        dp_issue_dw(tmem, tl, pl, C, 1, wi, true);
        lp_tc_commit(bar_dw);
      }
      lp_mbar_wait(bar, phase);
      phase ^= 1;
      lp_tc_fence_after();
      ++n_dw;
      LP_DP_LD(v);
#pragma unroll
      for (int j = 0; j < 32; ++j) S[j] = v[j];
    }
    if (me.active) {  // Hc == 0: S already is the gradient of (trunk + encoding) summed over the steps
      float4* ge = reinterpret_cast<float4*>(io.g_enc + (long long)me.ray * 32);
#pragma unroll
      for (int k = 0; k < 8; ++k) ge[k] = make_float4(S[4 * k], S[4 * k + 1], S[4 * k + 2], S[4 * k + 3]);
    }
  }
  if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  {
    for (int c = 0; c < pl.n_feat; ++c) lp_red_add1(io.g_params + pl.bc_off + c, c == 0 ? bl0 : (c == 1 ? bl1 : bl2));
    lp_red_add1(io.g_params + pl.bo_off, bl3);
    float v[32];
    const unsigned tlane = lp_taddr(tmem, warp, 0);
    if (warp == 0) {  // rows 0..31 of every window: gradient feature j = lane; columns: input features i
      for (int l = 0; l < pl.NL; ++l) {
        lp_tmem_ld32u(tlane + DT_ACC + 32 * l, v);
        for (int i = 0; i < pl.K[l]; ++i) lp_red_add1(io.g_params + pl.w_off[l] + i * 32 + lane, v[i]);
      }
      if (pl.Hc > 0) {
        lp_tmem_ld32u(tlane + DT_PE, v);
        for (int i = 0; i < 32; ++i) lp_red_add1(io.g_params + pl.w_off[c0] + i * 32 + lane, v[i]);
      }
    }
    for (int g = 0; g < 2; ++g) {  // bias gradients: row 32k + j of window g = layer 4g + k, feature j; column 0
      lp_tmem_ld32u(tlane + DT_ACCB + 16 * g, v);
      const int l = 4 * g + warp;
      if (l < pl.NL) lp_red_add1(io.g_params + pl.b_off[l] + lane, v[0]);
    }
    if (warp < 2) {  // output layers: rows 0..31 opacity input, 32..63 colour input; columns dlogit_0..2, g_raw
      lp_tmem_ld32u(tlane + DT_PL, v);
      if (warp == 0) lp_red_add1(io.g_params + pl.wo_off + lane * pl.wo_N, v[3]);
      else
        for (int c = 0; c < pl.n_feat; ++c) lp_red_add1(io.g_params + pl.wc_off + lane * pl.wc_N + c, v[c]);
    }
  }
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(tmem);
}
#undef LP_DP_ROUND
#undef LP_DP_LD

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
static inline bool lp_deep_plan(const LpRenderArgs& a, DeepPlan* pl) {
  const LpDecoder& D = a.D;
  if (D.use_color_grid) return false;
  const int Lt = D.trunk.n_layers, Lo = D.opacity.n_layers, Lc = D.color.n_layers;
  if (Lt < 1 || Lt > 4 || Lo < 1 || Lo > 4 || Lc < 1 || Lc > 4) return false;
  if (Lt == 2 && Lo == 2 && Lc == 2) return false;  // lp_render_tc.cuh
  if (Lt + Lo - 1 + Lc - 1 > DP_MAXL) return false;
  if ((D.C != 16 && D.C != 32) || D.n_feat > 3 || D.in_c != 32) return false;
  DeepPlan p;
  memset(&p, 0, sizeof(p));
  p.Lt = Lt; p.Ho = Lo - 1; p.Hc = Lc - 1; p.NL = Lt + p.Ho + p.Hc; p.n_feat = D.n_feat;
  int n = 0;
  for (int l = 0; l < Lt; ++l, ++n) {
    const LpLayer& L = D.trunk.l[l];
    if (L.N != 32 || L.K != (l == 0 ? D.C : 32) || !L.relu) return false;
    p.w_off[n] = L.w_off; p.b_off[n] = L.b_off; p.K[n] = L.K;
  }
  for (int l = 0; l < p.Ho; ++l, ++n) {
    const LpLayer& L = D.opacity.l[l];
    if (L.N != 32 || L.K != 32 || !L.relu) return false;
    p.w_off[n] = L.w_off; p.b_off[n] = L.b_off; p.K[n] = 32;
  }
  for (int l = 0; l < p.Hc; ++l, ++n) {
    const LpLayer& L = D.color.l[l];
    if (L.N != 32 || L.K != 32 || !L.relu) return false;
    p.w_off[n] = L.w_off; p.b_off[n] = L.b_off; p.K[n] = 32;
  }
  const LpLayer &lo = D.opacity.l[Lo - 1], &lc = D.color.l[Lc - 1];
  if (lo.K != 32 || lc.K != 32 || lo.relu || lc.relu) return false;
  p.wo_off = lo.w_off; p.wo_N = lo.N; p.bo_off = lo.b_off;
  p.wc_off = lc.w_off; p.wc_N = lc.N; p.bc_off = lc.b_off;
  long long elems = 0;
  for (int i = 0; i < a.G.n; ++i) elems = a.G.g[i].base + (long long)a.G.g[i].B * a.G.g[i].D * a.G.g[i].H * a.G.g[i].W * D.C;
  if (elems >= (1ll << 31)) return false;
  *pl = p;
  return true;
}
template <int C, bool SCAF>
static int lp_deep_render_forward_t(cudaStream_t st, const LpRenderArgs& a, const DeepPlan& pl, const float* params, float* out_len,
                                    float* out_nlt, float* out_feat, int feat_stride) {
  const int groups = 2;
  const size_t bytes = DpImg::FWD_END + 128;
  if (LP_TC_SET_SMEM((lp_render_fwd_deep_kernel<C, SCAF>), bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + GT - 1) / GT;
  int blocks = (tiles + groups - 1) / groups;
  if (blocks > lp_tc_num_sms()) blocks = lp_tc_num_sms();
  LP_LAUNCH((lp_render_fwd_deep_kernel<C, SCAF>), dim3(blocks), dim3(groups * GT), bytes, st, a.R, a.M, pl, a.G, a.SC, params, out_len,
            out_nlt, out_feat, feat_stride);
  return LP_OK;
}
static inline int lp_deep_render_forward(cudaStream_t st, const LpRenderArgs& a, const DeepPlan& pl, const float* params,
                                         float* out_len, float* out_nlt, float* out_feat, int feat_stride) {
  if (a.use_scaffold)
    return a.D.C == 16 ? lp_deep_render_forward_t<16, true>(st, a, pl, params, out_len, out_nlt, out_feat, feat_stride)
                       : lp_deep_render_forward_t<32, true>(st, a, pl, params, out_len, out_nlt, out_feat, feat_stride);
  return a.D.C == 16 ? lp_deep_render_forward_t<16, false>(st, a, pl, params, out_len, out_nlt, out_feat, feat_stride)
                     : lp_deep_render_forward_t<32, false>(st, a, pl, params, out_len, out_nlt, out_feat, feat_stride);
}
template <int C, bool SCAF>
static int lp_deep_render_backward_t(cudaStream_t st, const LpRenderArgs& a, const DeepPlan& pl, const float* params,
                                     const LpBwdIo& io) {
  const size_t bytes = DpImg::BYTES;
  if (LP_TC_SET_SMEM((lp_render_bwd_deep_kernel<C, SCAF>), bytes)) return LP_ERR_CUDA;
  int blocks = (a.R.n + GT - 1) / GT;
  if (blocks > lp_tc_num_sms()) blocks = lp_tc_num_sms();
  LP_LAUNCH((lp_render_bwd_deep_kernel<C, SCAF>), dim3(blocks), dim3(GT), bytes, st, a.R, a.M, pl, a.G, a.SC, params, io);
  return LP_OK;
}
static inline int lp_deep_render_backward(cudaStream_t st, const LpRenderArgs& a, const DeepPlan& pl, const float* params,
                                          const LpBwdIo& io) {
  if (a.use_scaffold)
    return a.D.C == 16 ? lp_deep_render_backward_t<16, true>(st, a, pl, params, io) : lp_deep_render_backward_t<32, true>(st, a, pl, params, io);
  return a.D.C == 16 ? lp_deep_render_backward_t<16, false>(st, a, pl, params, io) : lp_deep_render_backward_t<32, false>(st, a, pl, params, io);
}

}  // namespace lptc
