// Tensor-core Renderer FORWARD for hidden width 64 (the reference's example configuration,
// examples/config/synthetic_overfit.json: mlp_hidden_chn 64, 32-channel triplane, occupancy scaffold): same
// thread-per-sample tcgen05 scheme as lp_render_tc.cuh (read its header first), every per-sample row 64 wide:
//   t0  [C -> 64]   K = C,   N = 64        t1  [64 -> 64]  K = 64,  N = 64
//   o0 | c0         K = 64,  N = 128  (+ the ray-encoding product into the colour half, K = 64, N = 64)
// Tensor memory per group: A hi 0..31 / lo 32..63, encoding hi 64..95 / lo 96..127, D 128..255 -> two groups of 128
// threads per SM.  The backward of this shape still takes the generic kernels (its weight images + dW operand
// tiles need a 1-group, K-tiled variant of lp_render_bwd_tc_kernel: DESIGN.md section 5, "Next").
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

#define LP_W_ROUND(ISSUE)                  \
  lp_tmem_wait_st();                       \
  lp_tc_fence_before();                    \
  lp_bar_sync(1 + grp, GT);                \
  if (issuer) {                            \
    lp_tc_fence_after();                   \
    ISSUE;                                 \
    lp_tc_commit(bar);                     \
  }                                        \
  lp_mbar_wait(bar, phase);                \
  phase ^= 1;                              \
  lp_tc_fence_after();

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
    float nlt = 0.f, T = 1.f, acc_len = 0.f, acc_c[3] = {0.f, 0.f, 0.f};
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
      if (M.noise) raw += M.sigma * lp_sample_noise(M, me.ray, step);
      nlt += SCAF ? delta * M.gain * lp_softplus(raw) * occ : delta * M.gain * lp_softplus(raw);
      const float Tn = expf(-nlt);
      const float w = T - Tn;
      T = Tn;
      acc_len = fmaf(w, depth, acc_len);
      const float wc = SCAF ? w * occ : w;
      acc_c[0] = fmaf(wc, lp_sigmoid(lg0), acc_c[0]);
      acc_c[1] = fmaf(wc, lp_sigmoid(lg1), acc_c[1]);
      acc_c[2] = fmaf(wc, lp_sigmoid(lg2), acc_c[2]);
    }
    if (me.active) {
      out_len[me.ray] = acc_len;
      out_nlt[me.ray] = nlt;
      for (int c = 0; c < D.n_feat; ++c) out_feat[(long long)me.ray * feat_stride + c] = acc_c[c];
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

}  // namespace lptc
