// Renderer fast path, second generation: the per-sample MLP runs on the 5th-generation tensor cores
// (tcgen05) with one THREAD per sample.
//
// A group of 128 threads marches 128 rays in lock step.  At every step each thread gathers the grid
// features of its own sample into registers, splits them into two bf16 terms (x = hi + lo) and stores
// them as its row of the A operand in tensor memory (tcgen05.st).  One elected thread then issues the
// layer as M = 128 MMAs whose B operand (the weights, also hi + lo, built once per CTA) sits in shared
// memory; three products hi*Whi + lo*Whi + hi*Wlo give ~1e-5 relative accuracy with fp32 accumulation
// (tools/tc_test3.cu).  The accumulator comes back with tcgen05.ld as one 32-wide row per thread, so
// bias, ReLU, the next split and finally the 4-wide output layer and the compositing are plain
// per-thread code: no fragment layouts, no shuffles, no shared-memory transposes.  The colour branch's
// "trunk + ray encoding" input is fed as a K = 64 product [trunk | encoding] x [Wc0; Wc0], with the
// encoding staged in tensor memory once per ray; opacity and colour hidden layers share one N = 64 MMA.
// Several groups per CTA keep the tensor pipe and the issue slots busy while a group waits for its MMA.
//
// Reference semantics: lightplane/triton_src/templates/renderer_fw.py:85-375 (forward) with the MLP
// of triton_src/shared/fwbw_util.py:26-150; see DESIGN.md section 4.
#pragma once

#include "lp_render_fast.cuh"

namespace lptc {
using lpf::H;
using lpf::Ray1;
using lpf::Sched;

constexpr int GT = 128;  // threads = rays per group (the MMA's M)

// ---- shared-memory weight image (byte offsets).  bf16 tiles are K-major UMMA operands
// [n/8][k/8][8 n][8 k]; element (n, k) at ((n/8)*(K/8) + k/8)*128 + (n%8)*16 + (k%8)*2 ----
template <int C>
struct Img {
  static constexpr int T0_HI = 0;                  // [32 out][C in]
  static constexpr int T0_LO = T0_HI + 32 * C * 2;
  static constexpr int T1_HI = T0_LO + 32 * C * 2;  // [32][32]
  static constexpr int T1_LO = T1_HI + 2048;
  static constexpr int OC_HI = T1_LO + 2048;        // [64 out: opacity hidden | colour hidden][64 in: trunk | encoding]
  static constexpr int OC_LO = OC_HI + 8192;
  static constexpr int F32 = OC_LO + 8192;          // fp32: b_t0 b_t1 b_o0 b_c0 [4][32] | wo1[32] | Wc1[32][4] | b_last[4]
  static constexpr int FB = 0, FWO = 128, FWC = 160, FBL = 288, NF = 292;
  static constexpr int FWD_END = F32 + NF * 4;
};

LP_DEVICE unsigned short lp_bf16_rn(float x) { return (unsigned short)(lp_pack_bf16x2(x, 0.f) & 0xffffu); }
LP_DEVICE float lp_bf16_to_f(unsigned short h) { return __uint_as_float((unsigned)h << 16); }
LP_DEVICE void lp_put_w(unsigned char* sm, int hi_off, int lo_off, int n, int k, int K, float w) {
  const int idx = ((n >> 3) * (K >> 3) + (k >> 3)) * 64 + (n & 7) * 8 + (k & 7);
  const unsigned short h = lp_bf16_rn(w);
  reinterpret_cast<unsigned short*>(sm + hi_off)[idx] = h;
  reinterpret_cast<unsigned short*>(sm + lo_off)[idx] = lp_bf16_rn(w - lp_bf16_to_f(h));
}
template <int C>
LP_DEVICE void lp_build_img(unsigned char* sm, const float* __restrict__ P, const LpDecoder& D) {
  using I = Img<C>;
  const LpLayer &t0 = D.trunk.l[0], &t1 = D.trunk.l[1], &o0 = D.opacity.l[0], &o1 = D.opacity.l[1],
                &c0 = D.color.l[0], &c1 = D.color.l[1];
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int e = tid; e < 32 * C; e += nth) lp_put_w(sm, I::T0_HI, I::T0_LO, e & 31, e >> 5, C, P[t0.w_off + (e >> 5) * t0.N + (e & 31)]);
  for (int e = tid; e < 32 * 32; e += nth) lp_put_w(sm, I::T1_HI, I::T1_LO, e & 31, e >> 5, 32, P[t1.w_off + (e >> 5) * t1.N + (e & 31)]);
  for (int e = tid; e < 64 * 64; e += nth) {
    const int n = e & 63, k = e >> 6;
    float w;
    if (n < 32) w = k < 32 ? P[o0.w_off + k * o0.N + n] : 0.f;
    else w = P[c0.w_off + (k & 31) * c0.N + (n - 32)];
    lp_put_w(sm, I::OC_HI, I::OC_LO, n, k, 64, w);
  }
  float* F = reinterpret_cast<float*>(sm + I::F32);
  for (int e = tid; e < 32; e += nth) {
    F[I::FB + e] = P[t0.b_off + e];
    F[I::FB + 32 + e] = P[t1.b_off + e];
    F[I::FB + 64 + e] = P[o0.b_off + e];
    F[I::FB + 96 + e] = P[c0.b_off + e];
    F[I::FWO + e] = P[o1.w_off + e * o1.N];
    for (int c = 0; c < 4; ++c) F[I::FWC + 4 * e + c] = c < D.n_feat ? P[c1.w_off + e * c1.N + c] : 0.f;
  }
  if (tid < 4) F[I::FBL + tid] = tid == 3 ? P[o1.b_off] : (tid < D.n_feat ? P[c1.b_off + tid] : 0.f);
}

// x = hi + lo with hi = x truncated to bf16 and lo = bf16_rn(x - hi): two values -> one packed word each
LP_DEVICE void lp_split2(float x0, float x1, unsigned& hi, unsigned& lo) {
  hi = __byte_perm(__float_as_uint(x0), __float_as_uint(x1), 0x7632);
  lo = lp_pack_bf16x2(x0 - __uint_as_float(__float_as_uint(x0) & 0xffff0000u), x1 - __uint_as_float(__float_as_uint(x1) & 0xffff0000u));
}
// split a row of N values and store it as this thread's row of the A operand (hi at a_col, lo at a_col + 16)
template <int N>
LP_DEVICE void lp_stage_row(unsigned taddr_a, const float (&x)[N]) {
  unsigned hi[N / 2], lo[N / 2];
#pragma unroll
  for (int j = 0; j < N / 2; ++j) lp_split2(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
  lp_tmem_st<N / 2>(taddr_a, hi);
  lp_tmem_st<N / 2>(taddr_a + 16, lo);
}

// per-group tensor-memory columns (forward): A hi 0..15 / lo 16..31, encoding hi 32..47 / lo 48..63, D 64..127
constexpr int TC_A = 0, TC_E = 32, TC_D = 64, TC_GROUP_COLS = 128;

// leader thread: D(n columns) = A(K) x W, three bf16 products per 16-wide k-step
LP_DEVICE void lp_issue_layer(unsigned tbase, int d_col, int a_col, lp_kdesc_t whi, lp_kdesc_t wlo, int ksteps, int k0,
                              int nstride, int n, bool first) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    if (ks < ksteps) {
      const lp_kdesc_t bh = lp_tc_kadv(whi, (k0 + ks) * 256), bl = lp_tc_kadv(wlo, (k0 + ks) * 256);
      lp_tc_mma_ts(false, tbase + d_col, tbase + a_col + 8 * ks, bh, nstride, n, !(first && ks == 0));
      lp_tc_mma_ts(false, tbase + d_col, tbase + a_col + 16 + 8 * ks, bh, nstride, n, 1);
      lp_tc_mma_ts(false, tbase + d_col, tbase + a_col + 8 * ks, bl, nstride, n, 1);
    }
  }
}

// the owner thread samples all C channels of its sample point into registers
template <int C>
LP_DEVICE void lp_gather_regs(const LpGridSet& G, int b, float x, float y, float z, float oob, float (&acc)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  for (int gi = 0; gi < G.n; ++gi) {
    int off[8];
    float w[8];
    const int nt = lpf::lp_taps_i32(G.g[gi], C, b, x, y, z, off, w);
    float wsum = 0.f;
#pragma unroll
    for (int tp = 0; tp < 8; ++tp)
      if (tp < nt) wsum += w[tp];
    if (wsum == 0.f) continue;  // the sample misses this grid entirely
#pragma unroll
    for (int tp = 0; tp < 8; ++tp) {
      if (tp < nt) {
#pragma unroll
        for (int k = 0; k < C / 4; ++k) {
          const float4 v = lp_ldg4(G.data + off[tp] + 4 * k);
          acc[4 * k] = fmaf(w[tp], v.x, acc[4 * k]); acc[4 * k + 1] = fmaf(w[tp], v.y, acc[4 * k + 1]);
          acc[4 * k + 2] = fmaf(w[tp], v.z, acc[4 * k + 2]); acc[4 * k + 3] = fmaf(w[tp], v.w, acc[4 * k + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] *= oob;
}

// ===========================================================================================
// forward
// ===========================================================================================
template <int C>
__global__ void __launch_bounds__(512, 1) lp_render_fwd_tc_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G,
                                                                   const float* __restrict__ params,
                                                                   float* __restrict__ out_len, float* __restrict__ out_nlt,
                                                                   float* __restrict__ out_feat, int feat_stride) {
  using I = Img<C>;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid / GT, ngroups = blockDim.x / GT, wig = (tid >> 5) & 3;  // warp in group = TMEM lane quarter
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + I::FWD_END);  // one per group
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  lp_build_img<C>(sm, params, D);
  if (tid == 0) {
    for (int i = 0; i < ngroups; ++i) lp_mbar_init(bars + i, 1);
    lp_mbar_init_fence();
  }
  if (tid < 32) lp_tmem_alloc512(tmem_slot);
  lp_fence_async_smem();
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  const unsigned tbase = *tmem_slot + (unsigned)(grp * TC_GROUP_COLS);
  const unsigned tme = lp_taddr(tbase, wig, 0);  // this thread's lane, column 0 of the group
  const bool leader = (tid % GT) == 0;
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  const lp_kdesc_t w_t0h = lp_tc_kdesc_lo(sm + I::T0_HI), w_t0l = lp_tc_kdesc_lo(sm + I::T0_LO),
                   w_t1h = lp_tc_kdesc_lo(sm + I::T1_HI), w_t1l = lp_tc_kdesc_lo(sm + I::T1_LO),
                   w_och = lp_tc_kdesc_lo(sm + I::OC_HI), w_ocl = lp_tc_kdesc_lo(sm + I::OC_LO);
  unsigned long long* bar = bars + grp;
  int phase = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;

  for (int tile = blockIdx.x * ngroups + grp; tile < num_tiles; tile += gridDim.x * ngroups) {
    const Ray1 me = lpf::lp_load_ray1(R, tile * GT + (tid % GT), G.g[0].B);
    {  // stage the ray encoding once (columns TC_E..): A operand of the colour layer's second half
      float e[32];
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)(me.active ? me.ray : R.n - 1) * H);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 v = __ldg(e4 + k);
        e[4 * k] = v.x; e[4 * k + 1] = v.y; e[4 * k + 2] = v.z; e[4 * k + 3] = v.w;
      }
      lp_stage_row<32>(tme + TC_E, e);
    }
    float nlt = 0.f, T = 1.f, acc_len = 0.f, acc_c[3] = {0.f, 0.f, 0.f};

    for (int step = 0; step < tot; ++step) {
      const Sched sc = lpf::lp_sched(step, M);
      float depth, delta;
      lpf::lp_depth_delta(sc, me.near, me.far, depth, delta);
      {
        float x = me.ox + depth * me.dx, y = me.oy + depth * me.dy, z = me.oz + depth * me.dz;
        if (M.contract) lp_contract(x, y, z);
        const float oob = M.mask_oob ? lp_in_bounds(x, y, z) : 1.f;
        float x0[C];
        lp_gather_regs<C>(G, me.b, x, y, z, oob, x0);
        lp_stage_row<C>(tme + TC_A, x0);
      }
      // ---- trunk layer 0 ----
      lp_tmem_wait_st();
      lp_tc_fence_before();
      lp_bar_sync(1 + grp, GT);
      if (leader) {
        lp_tc_fence_after();
        lp_issue_layer(tbase, TC_D, TC_A, w_t0h, w_t0l, C / 16, 0, (C / 8) * 128, 32, true);
        lp_tc_commit(bar);
      }
      float v[32];
      lp_mbar_wait(bar, phase); phase ^= 1;
      lp_tc_fence_after();
      lp_tmem_ld32u(tme + TC_D, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + F[I::FB + j], 0.f);
      lp_stage_row<32>(tme + TC_A, v);
      // ---- trunk layer 1 ----
      lp_tmem_wait_st();
      lp_tc_fence_before();
      lp_bar_sync(1 + grp, GT);
      if (leader) {
        lp_tc_fence_after();
        lp_issue_layer(tbase, TC_D, TC_A, w_t1h, w_t1l, 2, 0, 512, 32, true);
        lp_tc_commit(bar);
      }
      lp_mbar_wait(bar, phase); phase ^= 1;
      lp_tc_fence_after();
      lp_tmem_ld32u(tme + TC_D, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + F[I::FB + 32 + j], 0.f);
      lp_stage_row<32>(tme + TC_A, v);
      // ---- opacity + colour hidden layers: [trunk | encoding] (K = 64) x [64 outputs] ----
      lp_tmem_wait_st();
      lp_tc_fence_before();
      lp_bar_sync(1 + grp, GT);
      if (leader) {
        lp_tc_fence_after();
        lp_issue_layer(tbase, TC_D, TC_A, w_och, w_ocl, 2, 0, 1024, 64, true);
        lp_issue_layer(tbase, TC_D, TC_E, w_och, w_ocl, 2, 2, 1024, 64, false);
        lp_tc_commit(bar);
      }
      lp_mbar_wait(bar, phase); phase ^= 1;
      lp_tc_fence_after();
      // ---- output layer (4 wide) on the CUDA cores, exact fp32 ----
      float raw = F[I::FBL + 3], lg0 = F[I::FBL], lg1 = F[I::FBL + 1], lg2 = F[I::FBL + 2];
      lp_tmem_ld32u(tme + TC_D, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) raw = fmaf(fmaxf(v[j] + F[I::FB + 64 + j], 0.f), F[I::FWO + j], raw);
      lp_tmem_ld32u(tme + TC_D + 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float hc = fmaxf(v[j] + F[I::FB + 96 + j], 0.f);
        const float4 w = *reinterpret_cast<const float4*>(F + I::FWC + 4 * j);
        lg0 = fmaf(hc, w.x, lg0); lg1 = fmaf(hc, w.y, lg1); lg2 = fmaf(hc, w.z, lg2);
      }
      // ---- compositing (renderer_fw.py:289-340) ----
      if (M.noise) raw += M.sigma * lp_sample_noise(M, me.ray, step);
      nlt += delta * M.gain * lp_softplus(raw);
      const float Tn = expf(-nlt);
      const float w = T - Tn;
      T = Tn;
      acc_len = fmaf(w, depth, acc_len);
      acc_c[0] = fmaf(w, lp_sigmoid(lg0), acc_c[0]);
      acc_c[1] = fmaf(w, lp_sigmoid(lg1), acc_c[1]);
      acc_c[2] = fmaf(w, lp_sigmoid(lg2), acc_c[2]);
    }
    if (me.active) {
      out_len[me.ray] = acc_len;
      out_nlt[me.ray] = nlt;
      for (int c = 0; c < D.n_feat; ++c) out_feat[(long long)me.ray * feat_stride + c] = acc_c[c];
    }
  }
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(*tmem_slot);
}

template <int C>
static int lp_tc_render_forward_t(cudaStream_t st, const LpRenderArgs& a, const float* params, float* out_len, float* out_nlt,
                                  float* out_feat, int feat_stride) {
  const int groups = 4;
  const size_t bytes = Img<C>::FWD_END + 128;
  if (LP_FAST_SET_SMEM(lp_render_fwd_tc_kernel<C>, bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + GT - 1) / GT;
  int blocks = (tiles + groups - 1) / groups;
  const int max_blocks = lp_fast_num_sms();  // persistent; one CTA per SM owns its tensor memory
  if (blocks > max_blocks) blocks = max_blocks;
  LP_LAUNCH(lp_render_fwd_tc_kernel<C>, dim3(blocks), dim3(groups * GT), bytes, st, a.R, a.M, a.D, a.G, params, out_len,
            out_nlt, out_feat, feat_stride);
  return LP_OK;
}
static inline int lp_tc_render_forward(cudaStream_t st, const LpRenderArgs& a, const float* params, float* out_len,
                                       float* out_nlt, float* out_feat, int feat_stride) {
  if (a.D.C == 16) return lp_tc_render_forward_t<16>(st, a, params, out_len, out_nlt, out_feat, feat_stride);
  return lp_tc_render_forward_t<32>(st, a, params, out_len, out_nlt, out_feat, feat_stride);
}

}  // namespace lptc
