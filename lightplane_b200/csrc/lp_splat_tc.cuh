// Tensor-core MLP splatter (reference: lightplane_splatter.py:167-338, splatter_fw.py:168-309,
// splatter_bw.py:183-394) for the shape  [c_in -> 32 -> c_out],  c_in, c_out in {16, 32}:
//   forward   x = sample(input_grid) + splatting_feature;  y = MLP(x);  splat y (and 1) into the output grid
//   backward  dY = sample(grad_output_grid) * valid;  MLP backward;  grad_feature += d_x;  splat d_x into the
//             input-grid gradient;  dW, db.
// Same thread-per-sample tcgen05 scheme as lp_render_tc.cuh (read its header first): 128 rays per group, A
// operand rows staged in tensor memory as bf16 hi + lo, weights as K-major hi + lo tiles in shared memory,
// parameter gradients as one [x | h | ones]^T x [d_h | dY] product per group-step accumulated in TMEM.
// Other MLP shapes take the generic kernels of lp_splat.cuh.
#pragma once

#include "lp_render_tc.cuh"
#include "lp_splat.cuh"

namespace lptc {

struct SpImg {  // byte offsets, sized for c_in = c_out = 32
  static constexpr int W0_HI = 0;                  // [32 hidden][c_in]
  static constexpr int W0_LO = W0_HI + 2048;
  static constexpr int W1_HI = W0_LO + 2048;       // [c_out][32 hidden]
  static constexpr int W1_LO = W1_HI + 2048;
  static constexpr int F32 = W1_LO + 2048;         // fp32: b0[32] b1[32]
  static constexpr int FWD_END = F32 + 256;
  static constexpr int X1_HI = FWD_END;            // d_h:  [32 hidden][c_out]
  static constexpr int X1_LO = X1_HI + 2048;
  static constexpr int X0_HI = X1_LO + 2048;       // d_x:  [c_in][32 hidden]
  static constexpr int X0_LO = X0_HI + 2048;
  static constexpr int BARS = X0_LO + 2048;
  static constexpr int GROUPS = BARS + 128;
  // per-group operand tiles (MN-major over the group's 128 samples, see lp_render_tc.cuh)
  static constexpr int A1 = 0;                      // rows: x (c_in) | h (32) | ones  -> at most 9 chunks
  static constexpr int DY = A1 + 9 * 2048;          // cols: d_h (32) | dY (c_out)      -> at most 8 chunks
  static constexpr int GROUP_BYTES = DY + 8 * 2048;
  static_assert(A1 + 16 * 2048 <= GROUP_BYTES, "operand window leaves the group's region");
};
constexpr int SP_A = 0, SP_D = 64, SP_GROUP_COLS = 96;  // TMEM per group: A hi 0..31 / lo 32..63, D 64..95
constexpr int SPB_W = 4 * SP_GROUP_COLS;                // CTA-wide dW accumulator (N = 32 + c_out <= 64)

LP_DEVICE void lp_build_spimg(unsigned char* sm, const float* __restrict__ P, const LpSplatMlp& S, bool with_dx) {
  using I = SpImg;
  const LpLayer &l0 = S.mlp.l[0], &l1 = S.mlp.l[1];
  const int tid = threadIdx.x, nth = blockDim.x, ci = S.c_in, co = S.c_out;
  for (int e = tid; e < 32 * ci; e += nth) {  // W0[k][n], k < c_in, n < 32
    const int n = e & 31, k = e >> 5;
    const float w = P[l0.w_off + k * l0.N + n];
    lp_put_w(sm, I::W0_HI, I::W0_LO, n, k, ci, w);
    if (with_dx) lp_put_w(sm, I::X0_HI, I::X0_LO, k, n, 32, w);
  }
  for (int e = tid; e < 32 * co; e += nth) {  // W1[k][n], k < 32, n < c_out
    const int n = e % co, k = e / co;
    const float w = P[l1.w_off + k * l1.N + n];
    lp_put_w(sm, I::W1_HI, I::W1_LO, n, k, 32, w);
    if (with_dx) lp_put_w(sm, I::X1_HI, I::X1_LO, k, n, co, w);
  }
  float* F = reinterpret_cast<float*>(sm + I::F32);
  for (int e = tid; e < 32; e += nth) {
    F[e] = P[l0.b_off + e];
    F[32 + e] = e < co ? P[l1.b_off + e] : 0.f;
  }
}

// scalar weight splat of one sample (the output grid's normaliser), 32-bit offsets
LP_DEVICE void lp_splat_weight_regs(const LpGridSet& G, float* weight, int b, float x, float y, float z, float scale) {
  if (weight == nullptr) return;
  for (int gi = 0; gi < G.n; ++gi) {
    int off[8];
    float w[8];
    const int nt = lp_taps_i32(G.g[gi], 1, b, x, y, z, off, w);  // C = 1: offsets in texels
#pragma unroll
    for (int tp = 0; tp < 8; ++tp)
      if (tp < nt && w[tp] != 0.f) lp_red_add1(weight + off[tp] - (int)G.g[gi].base + (int)(G.g[gi].base / G.C), w[tp] * scale);
  }
}

// ===========================================================================================
// forward
// ===========================================================================================
template <int CI, int CO>
__global__ void __launch_bounds__(512, 1) lp_mlp_splat_fwd_tc_kernel(LpRays R, LpMarch M, LpSplatMlp S, LpGridSet IN, LpGridSet OUT,
                                                                      float* __restrict__ weight, const float* __restrict__ valid,
                                                                      const float* __restrict__ params) {
  using I = SpImg;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x;
  const int grp = tid / GT, ngroups = blockDim.x / GT, wig = (tid >> 5) & 3;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + I::FWD_END);
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  lp_build_spimg(sm, params, S, false);
  if (tid == 0) {
    for (int i = 0; i < ngroups; ++i) lp_mbar_init(bars + i, 4);
    lp_mbar_init_fence();
  }
  if (tid < 32) lp_tmem_alloc512(tmem_slot);
  lp_fence_async_smem();
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  const unsigned tbase = *tmem_slot + (unsigned)(grp * SP_GROUP_COLS);
  const unsigned tme = lp_taddr(tbase, wig, 0);
  const bool issuer = (tid & 31) == 0;
  lp_tmem_zero<32>(tme + SP_D);
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  const lp_kdesc_t w0h = lp_tc_kdesc_lo(sm + I::W0_HI), w0l = lp_tc_kdesc_lo(sm + I::W0_LO),
                   w1h = lp_tc_kdesc_lo(sm + I::W1_HI), w1l = lp_tc_kdesc_lo(sm + I::W1_LO);
  unsigned long long* bar = bars + grp;
  int phase = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;

#define LP_SP_ROUND(ISSUE) LP_TCG_HANDOFF(1 + grp, GT, lp_elect_one(), ISSUE; lp_tc_commit(bar)) LP_TCG_WAIT(bar, phase)

  for (int tile = blockIdx.x * ngroups + grp; tile < num_tiles; tile += gridDim.x * ngroups) {
    const Ray1 me = lp_load_ray1(R, tile * GT + (tid % GT), OUT.g[0].B);
    const int q = me.active ? me.ray : R.n - 1;
    const float vm = me.active ? (valid ? valid[q] : 1.f) : 0.f;
    float feat[CI];
    {
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)q * CI);
#pragma unroll
      for (int k = 0; k < CI / 4; ++k) {
        const float4 v = __ldg(e4 + k);
        feat[4 * k] = v.x; feat[4 * k + 1] = v.y; feat[4 * k + 2] = v.z; feat[4 * k + 3] = v.w;
      }
    }
    for (int step = 0; step < tot; ++step) {
      const Sched sc = lp_sched(step, M);
      float depth, delta;
      lp_depth_delta(sc, me.near, me.far, depth, delta);
      float x = me.ox + depth * me.dx, y = me.oy + depth * me.dy, z = me.oz + depth * me.dz;
      if (M.contract) lp_contract(x, y, z);
      const float oob = M.mask_oob ? lp_in_bounds(x, y, z) : 1.f;
      {
        float xin[CI];
        lp_gather_regs<CI>(IN, me.b, x, y, z, oob, xin);
#pragma unroll
        for (int c = 0; c < CI; ++c) xin[c] += feat[c];
        lp_stage_row<CI, 32>(tme + SP_A, xin);
      }
      LP_SP_ROUND(lp_issue_layer_part(tbase, SP_D, SP_A, w0h, w0l, CI / 16, 0, (CI / 8) * 128, 32, 32, wig));
      float v[32];
      lp_tmem_ld32u(tme + SP_D, v);
      lp_tmem_zero<32>(tme + SP_D);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + F[j], 0.f);
      lp_stage_row<32, 32>(tme + SP_A, v);
      LP_SP_ROUND(lp_issue_layer_part(tbase, SP_D, SP_A, w1h, w1l, 2, 0, 512, CO, 32, wig));
      float yv[CO];
      lp_tmem_ld<CO>(tme + SP_D, yv);
      lp_tmem_zero<CO>(tme + SP_D);
      const float scale = oob * vm;
      if (scale != 0.f) {
#pragma unroll
        for (int c = 0; c < CO; ++c) yv[c] = (yv[c] + F[32 + c]) * scale;
        lp_splat_regs<CO>(OUT, OUT.data, me.b, x, y, z, yv);
        lp_splat_weight_regs(OUT, weight, me.b, x, y, z, scale);
      }
    }
  }
#undef LP_SP_ROUND
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(*tmem_slot);
}

// ===========================================================================================
// backward
// ===========================================================================================
template <int CI, int CO>
LP_DEVICE void lp_sp_issue_dw(unsigned tmem, unsigned char* gs, int accumulate, int wi) {
  using I = SpImg;
  const lp_kdesc_t a1 = lp_tc_mndesc_lo(gs + I::A1), dy = lp_tc_mndesc_lo(gs + I::DY);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    if (wi < 0 || (ks >> 1) == wi)
      lp_tc_mma_ss_mn(tmem + SPB_W, lp_tc_kadv(a1, ks * 256), lp_tc_kadv(dy, ks * 256), 2048, 32 + CO, accumulate | (ks > 0));
}

template <int CI, int CO>
__global__ void __launch_bounds__(512, 1) lp_mlp_splat_bwd_tc_kernel(LpRays R, LpMarch M, LpSplatMlp S, LpGridSet IN, LpGridSet GG,
                                                                      const float* __restrict__ valid,
                                                                      const float* __restrict__ params, float* __restrict__ g_feat,
                                                                      float* __restrict__ g_params, float* __restrict__ g_in) {
  using I = SpImg;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / GT, ngroups = blockDim.x / GT, s = tid % GT, wig = warp & 3;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + I::BARS);  // [2g] round trips, [2g+1] dW; [8] init
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 10);
  unsigned char* gs = sm + I::GROUPS + grp * I::GROUP_BYTES;
  lp_build_spimg(sm, params, S, true);
  for (int e = s; e < I::GROUP_BYTES / 16; e += GT) reinterpret_cast<uint4*>(gs)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  constexpr int ONES = CI / 8 + 4;  // A1 chunk whose first row is all ones (stack row CI + 32)
  *reinterpret_cast<unsigned short*>(gs + I::A1 + ONES * 2048 + (s >> 3) * 128 + (s & 7) * 16) = 0x3F80;
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
  if (tid == 0) {
    lp_sp_issue_dw<CI, CO>(tmem, gs, 0, -1);
    lp_tc_commit(bars + 8);
  }
  lp_mbar_wait(bars + 8, 0);
  lp_tc_fence_after();
  __syncthreads();

  const unsigned tbase = tmem + (unsigned)(grp * SP_GROUP_COLS);
  const unsigned tme = lp_taddr(tbase, wig, 0);
  const bool issuer = lane == 0;
  lp_tmem_zero<32>(tme + SP_D);
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  const lp_kdesc_t w0h = lp_tc_kdesc_lo(sm + I::W0_HI), w0l = lp_tc_kdesc_lo(sm + I::W0_LO),
                   x1h = lp_tc_kdesc_lo(sm + I::X1_HI), x1l = lp_tc_kdesc_lo(sm + I::X1_LO),
                   x0h = lp_tc_kdesc_lo(sm + I::X0_HI), x0l = lp_tc_kdesc_lo(sm + I::X0_LO);
  unsigned long long *bar = bars + 2 * grp, *bar_dw = bars + 2 * grp + 1;
  int phase = 0, n_dw = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;

#define LP_SP_ROUND(ISSUE) LP_TCG_HANDOFF(1 + grp, GT, lp_elect_one(), ISSUE) LP_TCG_WAIT(bar, phase)

  for (int tile = blockIdx.x * ngroups + grp; tile < num_tiles; tile += gridDim.x * ngroups) {
    const Ray1 me = lp_load_ray1(R, tile * GT + s, GG.g[0].B);
    const int q = me.active ? me.ray : R.n - 1;
    const float vm = me.active ? (valid ? valid[q] : 1.f) : 0.f;
    float feat[CI], gacc[CI];
    {
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)q * CI);
#pragma unroll
      for (int k = 0; k < CI / 4; ++k) {
        const float4 v = __ldg(e4 + k);
        feat[4 * k] = v.x; feat[4 * k + 1] = v.y; feat[4 * k + 2] = v.z; feat[4 * k + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < CI; ++c) gacc[c] = 0.f;
    }
    for (int step = 0; step < tot; ++step) {
      const Sched sc = lp_sched(step, M);
      float depth, delta;
      lp_depth_delta(sc, me.near, me.far, depth, delta);
      float x = me.ox + depth * me.dx, y = me.oy + depth * me.dy, z = me.oz + depth * me.dz;
      if (M.contract) lp_contract(x, y, z);
      const float oob = M.mask_oob ? lp_in_bounds(x, y, z) : 1.f;
      {
        float xin[CI];
        lp_gather_regs<CI>(IN, me.b, x, y, z, oob, xin);
#pragma unroll
        for (int c = 0; c < CI; ++c) xin[c] += feat[c];
        if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);  // previous step's dW products have consumed the tiles
        lp_tile_row<CI>(gs + I::A1, 0, s, xin);
        lp_stage_row<CI, 32>(tme + SP_A, xin);
      }
      LP_SP_ROUND(lp_issue_layer_part(tbase, SP_D, SP_A, w0h, w0l, CI / 16, 0, (CI / 8) * 128, 32, 32, wig); lp_tc_commit(bar));
      float v[32];
      lp_tmem_ld32u(tme + SP_D, v);
      lp_tmem_zero<32>(tme + SP_D);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + F[j], 0.f);
      lp_tile_row<32>(gs + I::A1, CI / 8, s, v);  // hidden activation: dW1 operand and ReLU gate
      {
        float dy[CO];  // upstream gradient of the MLP output = sample(grad_grid) * valid (splatter_bw.py:330-343)
        lp_gather_regs<CO>(GG, me.b, x, y, z, oob * vm, dy);
        lp_tile_row<CO>(gs + I::DY, 4, s, dy);
        lp_stage_row<CO, 32>(tme + SP_A, dy);
      }
      LP_SP_ROUND(lp_issue_layer_part(tbase, SP_D, SP_A, x1h, x1l, CO / 16, 0, (CO / 8) * 128, 32, 32, wig); lp_tc_commit(bar));
      lp_tmem_ld32u(tme + SP_D, v);
      lp_tmem_zero<32>(tme + SP_D);
      lp_gate_row<32>(v, gs + I::A1, CI / 8, s);  // d_h
      lp_tile_row<32>(gs + I::DY, 0, s, v);
      lp_stage_row<32, 32>(tme + SP_A, v);
      lp_fence_async_smem();
      LP_SP_ROUND(lp_issue_layer_part(tbase, SP_D, SP_A, x0h, x0l, 2, 0, 512, CI, 32, wig); lp_tc_commit(bar);
                  (lp_sp_issue_dw<CI, CO>(tmem, gs, 1, wig)); lp_tc_commit(bar_dw));
      ++n_dw;
      {
        float d[CI];
        lp_tmem_ld<CI>(tme + SP_D, d);
        lp_tmem_zero<CI>(tme + SP_D);
#pragma unroll
        for (int c = 0; c < CI; ++c) {
          gacc[c] += d[c];
          d[c] *= oob;
        }
        if (me.active && oob != 0.f) lp_splat_regs<CI>(IN, g_in, me.b, x, y, z, d);
      }
    }
    if (me.active) {
      float4* gf = reinterpret_cast<float4*>(g_feat + (long long)me.ray * CI);
#pragma unroll
      for (int k = 0; k < CI / 4; ++k) gf[k] = make_float4(gacc[4 * k], gacc[4 * k + 1], gacc[4 * k + 2], gacc[4 * k + 3]);
    }
  }
#undef LP_SP_ROUND
  if (n_dw > 0) lp_mbar_wait(bar_dw, (n_dw - 1) & 1);
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  if (warp < 4) {  // TMEM lane = stack row: x (CI) | h (32) | ones; columns: d_h (32) | dY (CO)
    const LpLayer &l0 = S.mlp.l[0], &l1 = S.mlp.l[1];
    float v[32];
    const unsigned tl = lp_taddr(tmem, warp, 0);
    const int row = 32 * warp + lane;
    lp_tmem_ld32u(tl + SPB_W, v);        // x d_h
    if (row < CI)
      for (int n = 0; n < 32; ++n) lp_red_add1(g_params + l0.w_off + row * l0.N + n, v[n]);
    if (row == CI + 32)
      for (int n = 0; n < 32; ++n) lp_red_add1(g_params + l0.b_off + n, v[n]);
    lp_tmem_ld32u(tl + SPB_W + 32, v);   // x dY
    if (row >= CI && row < CI + 32)
      for (int n = 0; n < CO; ++n) lp_red_add1(g_params + l1.w_off + (row - CI) * l1.N + n, v[n]);
    if (row == CI + 32)
      for (int n = 0; n < CO; ++n) lp_red_add1(g_params + l1.b_off + n, v[n]);
  }
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(tmem);
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
static inline bool lp_fits_i32(const LpGridSet& G) {
  long long elems = 0;
  for (int i = 0; i < G.n; ++i) elems = G.g[i].base + (long long)G.g[i].B * G.g[i].D * G.g[i].H * G.g[i].W * G.C;
  return elems < (1ll << 31);
}
static inline bool lp_tc_mlp_splat_supported(const LpSplatMlp& S, const LpGridSet& IN, const LpGridSet& OUT) {
  if (S.mlp.n_layers != 2) return false;
  if ((S.c_in != 16 && S.c_in != 32) || (S.c_out != 16 && S.c_out != 32)) return false;
  if (S.mlp.l[0].N != 32 || S.mlp.l[1].K != 32 || S.mlp.l[1].n_used != S.c_out) return false;
  if (IN.C != S.c_in || OUT.C != S.c_out) return false;
  return lp_fits_i32(IN) && lp_fits_i32(OUT);
}

template <int CI, int CO>
static int lp_tc_mlp_splat_forward_t(cudaStream_t st, const LpRays& R, const LpMarch& M, const LpSplatMlp& S, const LpGridSet& IN,
                                     const LpGridSet& O, float* weight, const float* valid, const float* params) {
  const int groups = 4;
  const size_t bytes = SpImg::FWD_END + 128;
  if (LP_TC_SET_SMEM((lp_mlp_splat_fwd_tc_kernel<CI, CO>), bytes)) return LP_ERR_CUDA;
  const int tiles = (R.n + GT - 1) / GT;
  int blocks = (tiles + groups - 1) / groups;
  if (blocks > lp_tc_num_sms()) blocks = lp_tc_num_sms();
  LP_LAUNCH((lp_mlp_splat_fwd_tc_kernel<CI, CO>), dim3(blocks), dim3(groups * GT), bytes, st, R, M, S, IN, O, weight, valid, params);
  return LP_OK;
}
template <int CI, int CO>
static int lp_tc_mlp_splat_backward_t(cudaStream_t st, const LpRays& R, const LpMarch& M, const LpSplatMlp& S, const LpGridSet& IN,
                                      const LpGridSet& GG, const float* valid, const float* params, float* g_feat, float* g_params,
                                      float* g_in) {
  const int groups = 4;
  const size_t bytes = SpImg::GROUPS + (size_t)groups * SpImg::GROUP_BYTES;
  if (LP_TC_SET_SMEM((lp_mlp_splat_bwd_tc_kernel<CI, CO>), bytes)) return LP_ERR_CUDA;
  const int tiles = (R.n + GT - 1) / GT;
  int blocks = (tiles + groups - 1) / groups;
  if (blocks > lp_tc_num_sms()) blocks = lp_tc_num_sms();
  LP_LAUNCH((lp_mlp_splat_bwd_tc_kernel<CI, CO>), dim3(blocks), dim3(groups * GT), bytes, st, R, M, S, IN, GG, valid, params, g_feat,
            g_params, g_in);
  return LP_OK;
}
#define LP_SP_DISPATCH(FN, ...)                                            \
  (S.c_in == 16 ? (S.c_out == 16 ? FN<16, 16>(__VA_ARGS__) : FN<16, 32>(__VA_ARGS__)) \
                : (S.c_out == 16 ? FN<32, 16>(__VA_ARGS__) : FN<32, 32>(__VA_ARGS__)))
static inline int lp_tc_mlp_splat_forward(cudaStream_t st, const LpRays& R, const LpMarch& M, const LpSplatMlp& S,
                                          const LpGridSet& IN, const LpGridSet& O, float* weight, const float* valid,
                                          const float* params) {
  return LP_SP_DISPATCH(lp_tc_mlp_splat_forward_t, st, R, M, S, IN, O, weight, valid, params);
}
static inline int lp_tc_mlp_splat_backward(cudaStream_t st, const LpRays& R, const LpMarch& M, const LpSplatMlp& S,
                                           const LpGridSet& IN, const LpGridSet& GG, const float* valid, const float* params,
                                           float* g_feat, float* g_params, float* g_in) {
  return LP_SP_DISPATCH(lp_tc_mlp_splat_backward_t, st, R, M, S, IN, GG, valid, params, g_feat, g_params, g_in);
}
#undef LP_SP_DISPATCH

}  // namespace lptc
