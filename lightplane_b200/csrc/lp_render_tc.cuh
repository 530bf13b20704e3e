// Renderer fast path for the default decoder shape (trunk/opacity/colour = 2/2/2 layers, hidden width
// 32, C in {16,32} grid channels, <= 3 colour channels, no separate colour grid): the per-sample MLP
// runs on the 5th-generation tensor cores (tcgen05) with one THREAD per sample.  Variants of the same scheme:
// lp_render_tc_cg.cuh (separate colour grid), lp_render_tc_wide.cuh (hidden width 64), lp_render_tc_deep.cuh (other
// layer counts), lp_splat_tc.cuh (MLP splatter); everything else takes the generic kernels of lp_render_generic.cuh.
//
// A group of 128 threads marches 128 rays in lock step.  At every step each thread gathers the grid
// features of its own sample into registers, splits them into two bf16 terms (x = hi + lo) and stores
// them as its row of the A operand in tensor memory (tcgen05.st).  Lane 0 of each of the group's warps then issues
// its share of the layer's M = 128 MMAs whose B operand (the weights, also hi + lo, built once per CTA) sits in shared
// memory; three products hi*Whi + lo*Whi + hi*Wlo give ~1e-5 relative accuracy with fp32 accumulation
// (tools/tc_test3.cu).  The accumulator comes back with tcgen05.ld as one 32-wide row per thread, so
// bias, ReLU, the next split and finally the 4-wide output layer and the compositing are plain
// per-thread code: no fragment layouts, no shuffles, no shared-memory transposes.  The colour branch's
// "trunk + ray encoding" input: the encoding's share enc x Wc0 + b is a per-ray constant, one product per ray tile,
// kept in shared memory and added where the bias would be; opacity and colour hidden layers share one N = 64 MMA.
// Several groups per CTA keep the tensor pipe and the issue slots busy while a group waits for its MMA.
// The epilogue arithmetic is written for issue slots: elect.sync issuers with warp-uniform operands,
// packed fp32 pairs, ReLU folded into the bf16 conversions, a triplane gather that shares the axes between planes.
//
// Reference semantics: lightplane/triton_src/templates/renderer_fw.py:85-375 (forward) with the MLP
// of triton_src/shared/fwbw_util.py:26-150; see DESIGN.md section 4.
#pragma once

#include "lp_render_generic.cuh"

#ifndef LP_TC_FWD_GROUPS
#define LP_TC_FWD_GROUPS 5  // groups of 128 threads per CTA in the forward kernel (TMEM: 96 columns each = 480 of 512; 96 registers
                            // per thread.  4 groups at 128 registers: 10.2 ms, 5 groups: 9.4 ms once the instruction count had come down)
#endif
#ifndef LP_TC_EMPTY_FOLD
#define LP_TC_EMPTY_FOLD 1  // reuse the decoder's zero-feature output at steps where a whole group is in empty space
#endif

namespace lptc {

constexpr int H = 32;  // hidden width of the decoder shape this path is specialised for

// ---- per-ray state, depth schedule and grid taps (semantics of lp_common.cuh, 32-bit offsets) ----
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
    // depth before the first background sample = the last regular depth: far, or near when S == 1 (naive_renderer.py:252-257)
    delta = depth - (s.first_inf ? (s.single ? near : (far - near) * 1.f + near) : far * s.prev);
  }
}

// ---- one tensor-core round trip of a group (shared by every kernel of the family) ----
// HANDOFF: publish this thread's staged operand row (tcgen05.wait::st + fence), meet the group at named barrier BARID,
// and let the issuing threads run ISSUE (which ends with their tcgen05.commit to the group's mbarrier);
// WAIT: block until the result is in tensor memory.  Independent work may sit between the two.
#define LP_TCG_HANDOFF(BARID, NTHREADS, ISSUER, ISSUE) \
  lp_tmem_wait_st();                                   \
  lp_tc_fence_before();                                \
  lp_bar_sync(BARID, NTHREADS);                        \
  if (ISSUER) {                                        \
    lp_tc_fence_after();                               \
    ISSUE;                                             \
  }
#define LP_TCG_WAIT(BAR, PHASE) \
  lp_mbar_wait(BAR, PHASE);     \
  PHASE ^= 1;                   \
  lp_tc_fence_after();

// ---- compositing of one sample, shared by every tensor-core renderer kernel ----
// forward (renderer_fw.py:289-340): NLT += delta*gain*softplus(raw)*occ; w = T_prev - T; len += w*depth; feat += w*occ*sigmoid(logit)
// Transcendentals of the compositing: ex2/lg2/rcp.approx forms (relative error ~2^-22, lp_platform.cuh).  The forward
// kernel and the backward kernel's recompute share them, so saved outputs and recomputed prefix sums stay consistent.
#ifndef LP_TC_FAST_MATH
#define LP_TC_FAST_MATH 1
#endif
// softplus(x) and its derivative sigmoid(x) from one exponential; log1p(e) = log(u) * e / (u - 1), u = 1 + e, keeps
// the relative accuracy for small e (func_util.py:19-28)
LP_DEVICE void lp_softplus_sig(float x, float& sp, float& sg) {
#if LP_TC_FAST_MATH
  const float e = lp_fast_exp(-fabsf(x));
  const float u = 1.f + e, d = u - 1.f, r = lp_fast_rcp(u);
  const float l1p = d == 0.f ? e : lp_fast_log(u) * e * lp_fast_rcp(d);
  sp = fmaxf(x, 0.f) + l1p;
  sg = x >= 0.f ? r : e * r;
#else
  sp = lp_softplus(x);
  sg = lp_sigmoid(x);
#endif
}
LP_DEVICE float lp_sig(float x) {
#if LP_TC_FAST_MATH
  return lp_fast_rcp(1.f + lp_fast_exp(-x));
#else
  return lp_sigmoid(x);
#endif
}
LP_DEVICE float lp_expneg(float x) {  // exp(-x)
#if LP_TC_FAST_MATH
  return lp_fast_exp(-x);
#else
  return expf(-x);
#endif
}
struct LpCompFwd {
  float nlt = 0.f, T = 1.f, len = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
  LP_DEVICE void add(const LpMarch& M, int ray, int step, float raw, float lg0, float lg1, float lg2, float depth, float delta, float occ) {
    if (M.noise) raw += M.sigma * lp_sample_noise(M, ray, step);
    float sp, sg;
    lp_softplus_sig(raw, sp, sg);
    nlt += delta * M.gain * sp * occ;
    const float Tn = lp_expneg(nlt);
    const float w = T - Tn;
    T = Tn;
    len = fmaf(w, depth, len);
    const float wc = w * occ;
    c0 = fmaf(wc, lp_sig(lg0), c0);
    c1 = fmaf(wc, lp_sig(lg1), c1);
    c2 = fmaf(wc, lp_sig(lg2), c2);
  }
};
// backward (renderer_bw.py:300-420), marching FORWARD with the saved outputs: with p_j = depth_j g_len + sum_c sigmoid_c gF_c,
// total = sum_j w_j p_j (from the saved outputs) and prefix_j = sum_{k<=j} w_k p_k,
//   dL/d(delta gain o_j) = T_j p_j - (total - prefix_j) + g_nlt      (the bracket is forced to 0 behind the last sample)
struct LpCompBwd {
  float g_len, g_nlt, gF[3], total;   // per-ray constants
  float nlt = 0.f, T = 1.f, prefix = 0.f;
  LP_DEVICE void init(const LpBwdIo& io, int q, bool active, int n_feat) {
    g_len = active ? io.g_len[q] : 0.f;
    g_nlt = active ? io.g_nlt[q] : 0.f;
    total = g_len * io.len[q];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      gF[c] = (active && c < n_feat) ? io.g_feat[(long long)q * io.g_feat_stride + c] : 0.f;
      if (c < n_feat) total = fmaf(gF[c], io.feat[(long long)q * io.feat_stride + c], total);
    }
    nlt = 0.f; T = 1.f; prefix = 0.f;
  }
  // returns the gradients of the raw opacity and of the three colour logits
  LP_DEVICE void grad(const LpMarch& M, int ray, int step, bool last, float raw, float lg0, float lg1, float lg2, float depth, float delta,
                      float occ, float& g_raw, float& dl0, float& dl1, float& dl2) {
    if (M.noise) raw += M.sigma * lp_sample_noise(M, ray, step);
    float sp, sg;
    lp_softplus_sig(raw, sp, sg);
    nlt += delta * M.gain * sp * occ;
    const float Tn = lp_expneg(nlt);
    const float w = T - Tn;
    T = Tn;
    const float s0 = lp_sig(lg0), s1 = lp_sig(lg1), s2 = lp_sig(lg2);
    const float p = fmaf(depth, g_len, fmaf(s0, gF[0], fmaf(s1, gF[1], s2 * gF[2]))) * occ;
    prefix = fmaf(w, p, prefix);
    const float suffix = last ? 0.f : total - prefix;
    const float g_dop = Tn * p - suffix + g_nlt;
    g_raw = g_dop * delta * M.gain * occ * sg;
    const float wo = w * occ;
    dl0 = wo * gF[0] * s0 * (1.f - s0);
    dl1 = wo * gF[1] * s1 * (1.f - s1);
    dl2 = wo * gF[2] * s2 * (1.f - s2);
  }
};

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
// `key` (optional): identity of the sample's footprint on this grid -- the index of its (unclamped) base cell in the grid
// padded by two cells per axis, batch included: two samples get the same taps iff their keys are equal
LP_DEVICE int lp_taps_i32(const LpGrid& g, int C, int b, float x, float y, float z, int* off, float* w, int* key = nullptr) {
  if (g.kind == LP_VOXEL) {
    int x0, y0, z0, cx[2], cy[2], cz[2];
    float fx, fy, fz, wx[2], wy[2], wz[2];
    lp_axis_i(x, g.W, x0, fx); lp_axis_i(y, g.H, y0, fy); lp_axis_i(z, g.D, z0, fz);
    if ((unsigned)(x0 + 1) > (unsigned)g.W || (unsigned)(y0 + 1) > (unsigned)g.H || (unsigned)(z0 + 1) > (unsigned)g.D)
      return 0;  // no corner inside the grid: every tap weight is zero
    if (key) *key = ((b * (g.D + 3) + z0 + 2) * (g.H + 3) + y0 + 2) * (g.W + 3) + x0 + 2;
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
  if ((unsigned)(u0 + 1) > (unsigned)U || (unsigned)(v0 + 1) > (unsigned)V) return 0;  // the sample misses this plane
  if (key) *key = (b * (V + 3) + v0 + 2) * (U + 3) + u0 + 2;
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

// ---- triplane fast path (LpGridSet::tri: the grids are exactly the XY, XZ and YZ planes of one W x H x D volume) ----
// The three planes share their axes: the index, fraction, clamped corners and weights of x, y and z are worked out ONCE
// per sample (the generic loop does it twice per plane, behind a run-time switch on the grid kind, before it can tell
// that a plane is missed), a plane is hit iff both of its axes are in range, and its taps are products / sums of the
// two axes' values -- the same numbers as lp_taps_i32 gives.
struct LpAxis {
  int i0, c0, c1;   // unclamped base cell, clamped corner indices
  float w0, w1;     // corner weights (0 outside the grid)
  bool ok;          // at least one corner inside
};
LP_DEVICE LpAxis lp_axis_pre(float p, int size) {
  LpAxis a;
  float fr;
  lp_axis_i(p, size, a.i0, fr);
  a.ok = (unsigned)(a.i0 + 1) <= (unsigned)size;
  lp_corner_i(a.i0, fr, size, a.w0, a.w1, a.c0, a.c1);
  return a;
}
LP_DEVICE LpAxis lp_axis_sel(bool first, const LpAxis& a, const LpAxis& b) {
  LpAxis r;
  r.i0 = first ? a.i0 : b.i0; r.c0 = first ? a.c0 : b.c0; r.c1 = first ? a.c1 : b.c1;
  r.w0 = first ? a.w0 : b.w0; r.w1 = first ? a.w1 : b.w1; r.ok = first ? a.ok : b.ok;
  return r;
}
// the four taps of a plane (fast axis u of size U, slow axis v) whose element offset starts at `bbase`
LP_DEVICE void lp_plane_taps_pre(int bbase, int U, int C, const LpAxis& au, const LpAxis& av, int* off, float* w) {
  off[0] = bbase + (av.c0 * U + au.c0) * C; w[0] = au.w0 * av.w0;
  off[1] = bbase + (av.c0 * U + au.c1) * C; w[1] = au.w1 * av.w0;
  off[2] = bbase + (av.c1 * U + au.c0) * C; w[2] = au.w0 * av.w1;
  off[3] = bbase + (av.c1 * U + au.c1) * C; w[3] = au.w1 * av.w1;
}

constexpr int GT = 128;  // threads = rays per group (the MMA's M)

// Ray handled by thread s of ray tile `tile`.  Default: 128 consecutive rays.  With the image-width hint
// (lp_march_cfg.ray_image_width) a tile is a 16x8-pixel block and a warp an 8x4 block of it, so that the samples a
// warp gathers / scatters at one step share texels (fewer distinct L1/L2 lines per request, fewer conflicting reductions).
LP_DEVICE int lp_tile_ray(const LpMarch& M, int tile, int s) {
  if (M.img_w <= 0) return tile * GT + s;
  const int tpr = M.img_w >> 4, ty = tile / tpr, tx = tile - ty * tpr;
  const int w = s >> 5, l = s & 31;
  return (ty * 8 + (w >> 1) * 4 + (l >> 3)) * M.img_w + tx * 16 + (w & 1) * 8 + (l & 7);
}

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
  static constexpr int F32 = OC_LO + 8192;          // fp32: b_t0 b_t1 b_o0 b_c0 [4][32] | wo1[32] | Wc1 as [16 row pairs][4 outputs][2 rows] | b_last[4]
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
    for (int c = 0; c < 4; ++c) F[I::FWC + 8 * (e >> 1) + 2 * c + (e & 1)] = c < D.n_feat ? P[c1.w_off + e * c1.N + c] : 0.f;  // pair layout, see Img
  }
  if (tid < 4) F[I::FBL + tid] = tid == 3 ? P[o1.b_off] : (tid < D.n_feat ? P[c1.b_off + tid] : 0.f);
}

// x = hi + lo with hi = x truncated to bf16 and lo = bf16_rn(x - hi): two values -> one packed word each
template <bool PK = true>
LP_DEVICE void lp_split2(float x0, float x1, unsigned& hi, unsigned& lo) {
  hi = __byte_perm(__float_as_uint(x0), __float_as_uint(x1), 0x7632);
  const float t0 = __uint_as_float(__float_as_uint(x0) & 0xffff0000u), t1 = __uint_as_float(__float_as_uint(x1) & 0xffff0000u);
  if constexpr (PK) {
    const float2 r = lp_sub2(lp_f2(x0, x1), lp_f2(t0, t1));
    lo = lp_pack_bf16x2(r.x, r.y);
  } else {
    lo = lp_pack_bf16x2(x0 - t0, x1 - t1);
  }
}
// split a row of N values and store it as this thread's row of the A operand (hi at a_col, lo at a_col + LO)
template <int N, int LO = 16, bool PK = true>
LP_DEVICE void lp_stage_row(unsigned taddr_a, const float (&x)[N]) {
  unsigned hi[N / 2], lo[N / 2];
#pragma unroll
  for (int j = 0; j < N / 2; ++j) lp_split2<PK>(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
  lp_tmem_st<N / 2>(taddr_a, hi);
  lp_tmem_st<N / 2>(taddr_a + LO, lo);
}

// The same with the layer's ReLU folded in (x = pre-activation): hi = max(trunc(x), 0) on the packed pair; the residual
// x - trunc(x) has the sign of x, so the ReLU of the low part's conversion zeroes it exactly when x <= 0.  Bit-identical to
// splitting max(x, 0), two instructions per pair cheaper.
template <bool PK = true>
LP_DEVICE void lp_split2_relu(float x0, float x1, unsigned& hi, unsigned& lo) {
  hi = lp_relu_bf16x2(__byte_perm(__float_as_uint(x0), __float_as_uint(x1), 0x7632));
  const float t0 = __uint_as_float(__float_as_uint(x0) & 0xffff0000u), t1 = __uint_as_float(__float_as_uint(x1) & 0xffff0000u);
  if constexpr (PK) {
    const float2 r = lp_sub2(lp_f2(x0, x1), lp_f2(t0, t1));
    lo = lp_pack_bf16x2_relu(r.x, r.y);
  } else {
    lo = lp_pack_bf16x2_relu(x0 - t0, x1 - t1);
  }
}
template <int N, int LO = 16, bool PK = true>
LP_DEVICE void lp_stage_row_relu(unsigned taddr_a, const float (&x)[N]) {
  unsigned hi[N / 2], lo[N / 2];
#pragma unroll
  for (int j = 0; j < N / 2; ++j) lp_split2_relu<PK>(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
  lp_tmem_st<N / 2>(taddr_a, hi);
  lp_tmem_st<N / 2>(taddr_a + LO, lo);
}
// v[j] += bias[j] (the ReLU follows in lp_stage_row_relu / lp_tile_row_relu)
template <int N, bool PK = true>
LP_DEVICE void lp_bias_add(float (&v)[N], const float* bias) {
#ifdef LP_ABL_NO_BIAS  // timing experiment only (wrong results): what preloading the bias into the accumulator could save
  return;
#endif
#pragma unroll
  for (int k = 0; k < N / 4; ++k) {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * k);
    if constexpr (PK) {
      const float2 lo = lp_add2(lp_f2(v[4 * k], v[4 * k + 1]), lp_f2(b.x, b.y)), hi = lp_add2(lp_f2(v[4 * k + 2], v[4 * k + 3]), lp_f2(b.z, b.w));
      v[4 * k] = lo.x; v[4 * k + 1] = lo.y; v[4 * k + 2] = hi.x; v[4 * k + 3] = hi.y;
    } else {
      v[4 * k] += b.x; v[4 * k + 1] += b.y; v[4 * k + 2] += b.z; v[4 * k + 3] += b.w;
    }
  }
}

// ---- decoder epilogue arithmetic on packed fp32 pairs (same roundings as the scalar forms; lp_platform.cuh) ----
// v[j] = max(v[j] + bias[j], 0) for a row of N accumulator values (bias in shared memory, 16-byte aligned)
template <int N, bool PK = true>
LP_DEVICE void lp_bias_relu(float (&v)[N], const float* bias) {
#pragma unroll
  for (int k = 0; k < N / 4; ++k) {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * k);
    if constexpr (!PK) {
      v[4 * k] = fmaxf(v[4 * k] + b.x, 0.f); v[4 * k + 1] = fmaxf(v[4 * k + 1] + b.y, 0.f);
      v[4 * k + 2] = fmaxf(v[4 * k + 2] + b.z, 0.f); v[4 * k + 3] = fmaxf(v[4 * k + 3] + b.w, 0.f);
      continue;
    }
    const float2 lo = lp_add2(lp_f2(v[4 * k], v[4 * k + 1]), lp_f2(b.x, b.y)), hi = lp_add2(lp_f2(v[4 * k + 2], v[4 * k + 3]), lp_f2(b.z, b.w));
    v[4 * k] = fmaxf(lo.x, 0.f); v[4 * k + 1] = fmaxf(lo.y, 0.f); v[4 * k + 2] = fmaxf(hi.x, 0.f); v[4 * k + 3] = fmaxf(hi.y, 0.f);
  }
}
// opacity head: bias + sum_j h[j] * wo[j], even and odd rows in the two halves of one packed accumulator
template <bool PK = true>
LP_DEVICE float lp_head_opacity(const float (&h)[32], const float* wo, float bias) {
  if constexpr (!PK) {
    float r0 = bias, r1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) { r0 = fmaf(h[j], wo[j], r0); r1 = fmaf(h[j + 1], wo[j + 1], r1); }
    return r0 + r1;
  }
  float2 r = lp_f2(bias, 0.f);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 w = *reinterpret_cast<const float4*>(wo + 4 * k);
    r = lp_fma2(lp_f2(h[4 * k], h[4 * k + 1]), lp_f2(w.x, w.y), r);
    r = lp_fma2(lp_f2(h[4 * k + 2], h[4 * k + 3]), lp_f2(w.z, w.w), r);
  }
  return r.x + r.y;
}
// colour head (3 outputs): weights in the image's pair layout [row pair][output][row parity]
template <bool PK = true>
LP_DEVICE void lp_head_colour(const float (&h)[32], const float* wc, const float* bias, float& lg0, float& lg1, float& lg2) {
  if constexpr (!PK) {
    float a0 = bias[0], a1 = bias[1], a2 = bias[2], b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      const float4 wa = *reinterpret_cast<const float4*>(wc + 8 * p), wb = *reinterpret_cast<const float4*>(wc + 8 * p + 4);
      a0 = fmaf(h[2 * p], wa.x, a0); a1 = fmaf(h[2 * p], wa.z, a1); a2 = fmaf(h[2 * p], wb.x, a2);
      b0 = fmaf(h[2 * p + 1], wa.y, b0); b1 = fmaf(h[2 * p + 1], wa.w, b1); b2 = fmaf(h[2 * p + 1], wb.y, b2);
    }
    lg0 = a0 + b0; lg1 = a1 + b1; lg2 = a2 + b2;
    return;
  }
  float2 a0 = lp_f2(bias[0], 0.f), a1 = lp_f2(bias[1], 0.f), a2 = lp_f2(bias[2], 0.f);
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float4 wa = *reinterpret_cast<const float4*>(wc + 8 * p), wb = *reinterpret_cast<const float4*>(wc + 8 * p + 4);
    const float2 hp = lp_f2(h[2 * p], h[2 * p + 1]);
    a0 = lp_fma2(hp, lp_f2(wa.x, wa.y), a0);
    a1 = lp_fma2(hp, lp_f2(wa.z, wa.w), a1);
    a2 = lp_fma2(hp, lp_f2(wb.x, wb.y), a2);
  }
  lg0 = a0.x + a0.y; lg1 = a1.x + a1.y; lg2 = a2.x + a2.y;
}
// backward of the two heads: d_ho[j] = g_raw * wo[j];  d_hc[j] = dl0 * Wc[j][0] + dl1 * Wc[j][1] + dl2 * Wc[j][2]
template <bool PK = true>
LP_DEVICE void lp_head_opacity_bwd(float (&d)[32], const float* wo, float g_raw) {
  if constexpr (!PK) {
#pragma unroll
    for (int j = 0; j < 32; ++j) d[j] = g_raw * wo[j];
    return;
  }
  const float2 g = lp_f2(g_raw, g_raw);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 w = *reinterpret_cast<const float4*>(wo + 4 * k);
    const float2 lo = lp_mul2(g, lp_f2(w.x, w.y)), hi = lp_mul2(g, lp_f2(w.z, w.w));
    d[4 * k] = lo.x; d[4 * k + 1] = lo.y; d[4 * k + 2] = hi.x; d[4 * k + 3] = hi.y;
  }
}
template <bool PK = true>
LP_DEVICE void lp_head_colour_bwd(float (&d)[32], const float* wc, float dl0, float dl1, float dl2) {
  if constexpr (!PK) {
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      const float4 wa = *reinterpret_cast<const float4*>(wc + 8 * p), wb = *reinterpret_cast<const float4*>(wc + 8 * p + 4);
      d[2 * p] = fmaf(dl0, wa.x, fmaf(dl1, wa.z, dl2 * wb.x));
      d[2 * p + 1] = fmaf(dl0, wa.y, fmaf(dl1, wa.w, dl2 * wb.y));
    }
    return;
  }
  const float2 g0 = lp_f2(dl0, dl0), g1 = lp_f2(dl1, dl1), g2 = lp_f2(dl2, dl2);
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float4 wa = *reinterpret_cast<const float4*>(wc + 8 * p), wb = *reinterpret_cast<const float4*>(wc + 8 * p + 4);
    const float2 r = lp_fma2(g0, lp_f2(wa.x, wa.y), lp_fma2(g1, lp_f2(wa.z, wa.w), lp_mul2(g2, lp_f2(wb.x, wb.y))));
    d[2 * p] = r.x; d[2 * p + 1] = r.y;
  }
}

// per-group tensor-memory columns (forward): A hi 0..15 / lo 16..31, D 32..95
constexpr int TC_A = 0, TC_D = 32, TC_GROUP_COLS = 96;  // (up to five groups fit the SM's 512 columns)

// single-issuer form (LP_TC_ISSUERS == 1 builds): D(n columns) = A(K) x W, three bf16 products per 16-wide k-step
LP_DEVICE void lp_issue_layer(unsigned tbase, int d_col, int a_col, lp_kdesc_t whi, lp_kdesc_t wlo, int ksteps, int k0,
                              int nstride, int n, bool first, int lo_off = 16) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    if (ks < ksteps) {
      const lp_kdesc_t bh = lp_tc_kadv(whi, (k0 + ks) * 256), bl = lp_tc_kadv(wlo, (k0 + ks) * 256);
      lp_tc_mma_ts(false, tbase + d_col, tbase + a_col + 8 * ks, bh, nstride, n, !(first && ks == 0));
      lp_tc_mma_ts(false, tbase + d_col, tbase + a_col + lo_off + 8 * ks, bh, nstride, n, 1);
      lp_tc_mma_ts(false, tbase + d_col, tbase + a_col + 8 * ks, bl, nstride, n, 1);
    }
  }
}

// Multi-issuer form used by the backward kernel: lane 0 of the group's warps 0..3 is issuer wi = 0..3 and
// takes k-step wi of the product (three MMAs); every MMA accumulates, the epilogue threads having cleared
// the accumulator (lp_tmem_zero) after reading the previous result, so the issue order is irrelevant.
LP_DEVICE void lp_issue_layer_part(unsigned tbase, int d_col, int a_col, lp_kdesc_t whi, lp_kdesc_t wlo, int ksteps, int k0,
                                   int nstride, int n, int lo_off, int wi) {
  if (wi >= 0 && wi < ksteps) {
    const lp_kdesc_t bh = lp_tc_kadv(whi, (k0 + wi) * 256), bl = lp_tc_kadv(wlo, (k0 + wi) * 256);
    lp_tc_mma_ts(false, tbase + d_col, tbase + a_col + 8 * wi, bh, nstride, n, 1);
    lp_tc_mma_ts(false, tbase + d_col, tbase + a_col + lo_off + 8 * wi, bh, nstride, n, 1);
    lp_tc_mma_ts(false, tbase + d_col, tbase + a_col + 8 * wi, bl, nstride, n, 1);
  }
}

// the owner thread samples all C channels of its sample point into registers
// (CW channels starting at ch0: a sample's channels may be split over several threads)
// Returns whether the sample touches any grid at all (false: acc is exactly zero).
template <int C, int CW = C, bool TRI = true>
LP_DEVICE bool lp_gather_regs(const LpGridSet& G, int b, float x, float y, float z, float oob, float (&acc)[CW], int ch0 = 0) {
  bool hit = false;
#pragma unroll
  for (int c = 0; c < CW; ++c) acc[c] = 0.f;
#ifdef LP_ABL_NO_MEM
  acc[0] = x * y + z;
  return true;
#endif
  if (TRI && G.tri) {  // XY, XZ, YZ planes of one volume (lp_cabi.cu lp_make_gridset): axes once, planes unrolled
    const int W = G.g[0].W, Hh = G.g[0].H, Dd = G.g[1].D;
    const LpAxis ax = lp_axis_pre(x, W), ay = lp_axis_pre(y, Hh), az = lp_axis_pre(z, Dd);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const LpAxis& au = p == 2 ? ay : ax;
      const LpAxis& av = p == 0 ? ay : az;
      if (!(au.ok && av.ok)) continue;  // the sample misses this plane
      hit = true;
      const int U = p == 2 ? Hh : W, V = p == 0 ? Hh : Dd;
      int off[4];
      float w[4];
      lp_plane_taps_pre((int)G.g[p].base + b * U * V * C, U, C, au, av, off, w);
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {
        const float2 wp = lp_f2(w[tp], w[tp]);
#pragma unroll
        for (int k = 0; k < CW / 4; ++k) {
          const float4 v = lp_ldg4(G.data + off[tp] + ch0 + 4 * k);
          const float2 lo = lp_fma2(wp, lp_f2(v.x, v.y), lp_f2(acc[4 * k], acc[4 * k + 1])), hi = lp_fma2(wp, lp_f2(v.z, v.w), lp_f2(acc[4 * k + 2], acc[4 * k + 3]));
          acc[4 * k] = lo.x; acc[4 * k + 1] = lo.y; acc[4 * k + 2] = hi.x; acc[4 * k + 3] = hi.y;
        }
      }
    }
    const float2 ob = lp_f2(oob, oob);
#pragma unroll
    for (int c = 0; c < CW; c += 2) {
      const float2 t = lp_mul2(lp_f2(acc[c], acc[c + 1]), ob);
      acc[c] = t.x; acc[c + 1] = t.y;
    }
    return hit && oob != 0.f;
  }
  for (int gi = 0; gi < G.n; ++gi) {
    int off[8];
    float w[8];
    const int nt = lp_taps_i32(G.g[gi], C, b, x, y, z, off, w);
    if (nt == 0) continue;  // the sample misses this grid entirely
    hit = true;
#pragma unroll
    for (int tp = 0; tp < 8; ++tp) {
      if (tp < nt) {
#pragma unroll
        for (int k = 0; k < CW / 4; ++k) {
          const float4 v = lp_ldg4(G.data + off[tp] + ch0 + 4 * k);
          acc[4 * k] = fmaf(w[tp], v.x, acc[4 * k]); acc[4 * k + 1] = fmaf(w[tp], v.y, acc[4 * k + 1]);
          acc[4 * k + 2] = fmaf(w[tp], v.z, acc[4 * k + 2]); acc[4 * k + 3] = fmaf(w[tp], v.w, acc[4 * k + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CW; ++c) acc[c] *= oob;
  return hit && oob != 0.f;
}

// ===========================================================================================
// forward
// ===========================================================================================
template <int C, bool SCAF>
__global__ void __launch_bounds__(LP_TC_FWD_GROUPS * 128, 1) lp_render_fwd_tc_kernel(LpRays R, LpMarch M, LpDecoder D, LpGridSet G,
                                                                   LpGridSet SC,
                                                                   const float* __restrict__ params,
                                                                   float* __restrict__ out_len, float* __restrict__ out_nlt,
                                                                   float* __restrict__ out_feat, int feat_stride) {
  using I = Img<C>;
  LP_DYN_SMEM(unsigned char, sm);
  const int tid = threadIdx.x;
  // warp index via a broadcast shuffle: what derives from it (group, tensor-/shared-memory operand addresses, mbarrier)
  // is then known to be warp-uniform and stays on the uniform datapath (see lp_render_tc_bwd.cuh)
  const int warp_u = LP_WARP_UNIFORM(tid >> 5);
  const int grp = warp_u >> 2, ngroups = blockDim.x / GT, wig = warp_u & 3;  // warp in group = TMEM lane quarter
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + I::FWD_END);  // one per group
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 8);
  // The CTA's ray tiles are handed to its groups from a shared-memory counter: tiles differ in how many steps fold, a
  // fixed stride per group leaves groups idle at the end (forward 9.41 -> 9.25 ms; across CTAs the split stays static).
  int* tile_ctr = reinterpret_cast<int*>(bars + 9);
  volatile int* tile_slot = reinterpret_cast<volatile int*>(bars + 10) + grp;
  lp_build_img<C>(sm, params, D);
  if (tid == 0) {
    for (int i = 0; i < ngroups; ++i) lp_mbar_init(bars + i, 4);  // four issuing threads per group
    lp_mbar_init_fence();
    *tile_ctr = 0;
  }
  if (tid < 32) lp_tmem_alloc512(tmem_slot);
  lp_fence_async_smem();
  lp_tc_fence_before();
  __syncthreads();
  lp_tc_fence_after();
  const unsigned tbase = LP_WARP_UNIFORM(*tmem_slot) + (unsigned)(grp * TC_GROUP_COLS);
  const unsigned tme = lp_taddr(tbase, wig, 0);  // this thread's lane, column 0 of the group
  // one elected lane (elect.sync: the compiler then emits the tcgen05 issue without an election loop) of each of the
  // group's four warps issues k-step `wig` (lp_issue_layer_part)
  lp_tmem_zero<32>(tme + TC_D);
  lp_tmem_zero<32>(tme + TC_D + 32);
  const float* F = reinterpret_cast<const float*>(sm + I::F32);
  const lp_kdesc_t w_t0h = lp_tc_kdesc_lo(sm + I::T0_HI), w_t0l = lp_tc_kdesc_lo(sm + I::T0_LO),
                   w_t1h = lp_tc_kdesc_lo(sm + I::T1_HI), w_t1l = lp_tc_kdesc_lo(sm + I::T1_LO),
                   w_och = lp_tc_kdesc_lo(sm + I::OC_HI), w_ocl = lp_tc_kdesc_lo(sm + I::OC_LO);
  unsigned long long* bar = bars + grp;
  float4* ecb = reinterpret_cast<float4*>(sm + I::FWD_END + 128 + grp * 16384) + (tid % GT);  // float4 [8][128] per group
  int phase = 0;
  const int num_tiles = (R.n + GT - 1) / GT;
  const int tot = M.S + M.S_inf;

  for (;;) {
    if (tid % GT == 0) *tile_slot = atomicAdd(tile_ctr, 1);
    lp_bar_sync(1 + grp, GT);
    // CTA b owns tiles b, b + gridDim.x, ...: a spread over the image (runs of neighbouring tiles per CTA measured slower: 9.64 ms);
    // the slot is rewritten only after this tile's many group barriers
    const int tile = blockIdx.x + *tile_slot * gridDim.x;
    if (tile >= num_tiles) break;
    const Ray1 me = lp_load_ray1(R, lp_tile_ray(M, tile, tid % GT), G.g[0].B);
    {  // The ray encoding's share of the colour hidden layer, enc x Wc0 + b, is a per-ray constant: one product per ray
       // tile (k-steps 2, 3 of the [opacity | colour] tile = the rows the encoding meets), kept in shared memory,
       // where it stands in for the bias.  (Same arithmetic, in the same order, as the backward kernel's recompute.)
      float e[32];
      const float4* e4 = reinterpret_cast<const float4*>(R.enc + (long long)(me.active ? me.ray : R.n - 1) * H);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 v = __ldg(e4 + k);
        e[4 * k] = v.x; e[4 * k + 1] = v.y; e[4 * k + 2] = v.z; e[4 * k + 3] = v.w;
      }
      lp_stage_row<32>(tme + TC_A, e);
      LP_TCG_HANDOFF(1 + grp, GT, lp_elect_one(), lp_issue_layer_part(tbase, TC_D, TC_A, w_och, w_ocl, 2, 2, 1024, 64, 16, wig); lp_tc_commit(bar));
      LP_TCG_WAIT(bar, phase);
      lp_tmem_ld32u(tme + TC_D + 32, e);
      lp_tmem_zero<32>(tme + TC_D + 32);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        ecb[k * GT] = make_float4(e[4 * k] + F[I::FB + 96 + 4 * k], e[4 * k + 1] + F[I::FB + 97 + 4 * k],
                                  e[4 * k + 2] + F[I::FB + 98 + 4 * k], e[4 * k + 3] + F[I::FB + 99 + 4 * k]);
    }
    LpCompFwd cf;
    // Empty-space folding.  A sample that misses every grid (or is masked out of bounds) has all-zero features,
    // so the decoder's output there does not depend on the position: it is evaluated once per ray (iteration
    // step = -1, "probe") and steps at which ALL 128 samples of the group are empty reuse it and skip the three
    // tensor-core round trips.  The vote rides on the first hand-off's barrier, so other steps pay nothing.
    float e_raw = 0.f, e_lg0 = 0.f, e_lg1 = 0.f, e_lg2 = 0.f;

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
          if (SCAF) {  // occupancy scaffold (renderer_fw.py:234-252): a step whose 128 samples are all in
            occ = lp_nearest(SC, me.b, x, y, z);  // empty space changes nothing and is skipped by the whole group
            if (!lp_bar_any(1 + grp, GT, occ != 0.f)) continue;
          }
          const float oob = M.mask_oob ? lp_in_bounds(x, y, z) : 1.f;
          hit = lp_gather_regs<C>(G, me.b, x, y, z, oob, x0);
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) x0[c] = 0.f;
        }
        lp_stage_row<C>(tme + TC_A, x0);
      }
      float raw, lg0, lg1, lg2;
      // ---- trunk layer 0 ----
      lp_tmem_wait_st();
      lp_tc_fence_before();
      const bool full = LP_TC_EMPTY_FOLD ? (lp_bar_any(1 + grp, GT, hit) || probe) : (lp_bar_sync(1 + grp, GT), true);
      if (full) {
      if (lp_elect_one()) {
        lp_tc_fence_after();
        lp_issue_layer_part(tbase, TC_D, TC_A, w_t0h, w_t0l, C / 16, 0, (C / 8) * 128, 32, 16, wig);
        lp_tc_commit(bar);
      }
      float v[32];
      lp_mbar_wait(bar, phase); phase ^= 1;
      lp_tc_fence_after();
      lp_tmem_ld32u(tme + TC_D, v);
      lp_tmem_zero<32>(tme + TC_D);
      lp_bias_add<32>(v, F + I::FB);
      lp_stage_row_relu<32>(tme + TC_A, v);
      // ---- trunk layer 1 ----
      lp_tmem_wait_st();
      lp_tc_fence_before();
      lp_bar_sync(1 + grp, GT);
      if (lp_elect_one()) {
        lp_tc_fence_after();
        lp_issue_layer_part(tbase, TC_D, TC_A, w_t1h, w_t1l, 2, 0, 512, 32, 16, wig);
        lp_tc_commit(bar);
      }
      lp_mbar_wait(bar, phase); phase ^= 1;
      lp_tc_fence_after();
      lp_tmem_ld32u(tme + TC_D, v);
      lp_tmem_zero<32>(tme + TC_D);
      lp_bias_add<32>(v, F + I::FB + 32);
      lp_stage_row_relu<32>(tme + TC_A, v);
      // ---- opacity + colour hidden layers: trunk (K = 32) x [64 outputs]; the encoding's share comes from `ecb` ----
      lp_tmem_wait_st();
      lp_tc_fence_before();
      lp_bar_sync(1 + grp, GT);
      if (lp_elect_one()) {
        lp_tc_fence_after();
        lp_issue_layer_part(tbase, TC_D, TC_A, w_och, w_ocl, 2, 0, 1024, 64, 16, wig);
        lp_tc_commit(bar);
      }
      lp_mbar_wait(bar, phase); phase ^= 1;
      lp_tc_fence_after();
      // ---- output layer (4 wide) on the CUDA cores, exact fp32 (two partial sums per output, as in the backward) ----
      {
        lp_tmem_ld32u(tme + TC_D, v);
        lp_tmem_zero<32>(tme + TC_D);
        lp_bias_relu<32>(v, F + I::FB + 64);
        raw = lp_head_opacity(v, F + I::FWO, F[I::FBL + 3]);
        lp_tmem_ld32u(tme + TC_D + 32, v);
        lp_tmem_zero<32>(tme + TC_D + 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // enc x Wc0 + b (per ray) stands in for the bias
          const float4 eb = ecb[k * GT];
          const float2 lo = lp_add2(lp_f2(v[4 * k], v[4 * k + 1]), lp_f2(eb.x, eb.y)), hi = lp_add2(lp_f2(v[4 * k + 2], v[4 * k + 3]), lp_f2(eb.z, eb.w));
          v[4 * k] = fmaxf(lo.x, 0.f); v[4 * k + 1] = fmaxf(lo.y, 0.f); v[4 * k + 2] = fmaxf(hi.x, 0.f); v[4 * k + 3] = fmaxf(hi.y, 0.f);
        }
        lp_head_colour(v, F + I::FWC, F + I::FBL, lg0, lg1, lg2);
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
  lp_tc_fence_before();
  __syncthreads();
  if (tid < 32) lp_tmem_dealloc512(*tmem_slot);
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
static inline bool lp_tc_render_supported(const LpRenderArgs& a) {
  const LpDecoder& D = a.D;
  if (D.use_color_grid) return false;
  if (D.trunk.n_layers != 2 || D.opacity.n_layers != 2 || D.color.n_layers != 2) return false;
  if (D.C != 16 && D.C != 32) return false;
  if (D.n_feat > 3 || D.in_c != lptc::H) return false;
  const LpLayer* ls[4] = {&D.trunk.l[0], &D.trunk.l[1], &D.opacity.l[0], &D.color.l[0]};
  for (int i = 0; i < 4; ++i)
    if (ls[i]->N != lptc::H) return false;
  long long elems = 0;  // the kernels address the grid with 32-bit element offsets
  for (int i = 0; i < a.G.n; ++i) {
    const LpGrid& g = a.G.g[i];
    elems = g.base + (long long)g.B * g.D * g.H * g.W * D.C;
  }
  if (elems >= (1ll << 31)) return false;
  return true;
}

#ifdef LP_HOSTSIM
#define LP_TC_SET_SMEM(kernel, bytes) 0
static inline int lp_tc_num_sms() { return 2; }
#else
#define LP_TC_SET_SMEM(kernel, bytes) \
  (cudaFuncSetAttribute((kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)) != cudaSuccess)
static inline int lp_tc_num_sms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
#endif

template <int C, bool SCAF>
static int lp_tc_render_forward_t(cudaStream_t st, const LpRenderArgs& a, const float* params, float* out_len, float* out_nlt,
                                  float* out_feat, int feat_stride) {
  const int groups = LP_TC_FWD_GROUPS;
  const size_t bytes = Img<C>::FWD_END + 128 + (size_t)groups * 16384;  // weights, mbarriers, per group the rays' enc x Wc0 + b
  if (LP_TC_SET_SMEM((lp_render_fwd_tc_kernel<C, SCAF>), bytes)) return LP_ERR_CUDA;
  const int tiles = (a.R.n + GT - 1) / GT;
  int blocks = (tiles + groups - 1) / groups;
  const int max_blocks = lp_tc_num_sms();  // persistent; one CTA per SM owns its tensor memory
  if (blocks > max_blocks) blocks = max_blocks;
  LP_LAUNCH((lp_render_fwd_tc_kernel<C, SCAF>), dim3(blocks), dim3(groups * GT), bytes, st, a.R, a.M, a.D, a.G, a.SC, params, out_len,
            out_nlt, out_feat, feat_stride);
  return LP_OK;
}
static inline int lp_tc_render_forward(cudaStream_t st, const LpRenderArgs& a, const float* params, float* out_len,
                                       float* out_nlt, float* out_feat, int feat_stride) {
  if (a.use_scaffold)
    return a.D.C == 16 ? lp_tc_render_forward_t<16, true>(st, a, params, out_len, out_nlt, out_feat, feat_stride)
                       : lp_tc_render_forward_t<32, true>(st, a, params, out_len, out_nlt, out_feat, feat_stride);
  return a.D.C == 16 ? lp_tc_render_forward_t<16, false>(st, a, params, out_len, out_nlt, out_feat, feat_stride)
                     : lp_tc_render_forward_t<32, false>(st, a, params, out_len, out_nlt, out_feat, feat_stride);
}


// ===========================================================================================
// backward: helpers shared by every tensor-core backward kernel (the default-shape kernel itself is lp_render_tc_bwd.cuh)
// ===========================================================================================
// this thread's sample s: features 8*chunk .. 8*chunk+7 of a tile
LP_DEVICE void lp_tile8(unsigned char* tile, int chunk, int s, float x0, float x1, float x2, float x3, float x4, float x5,
                        float x6, float x7) {
#ifdef LP_ABL_NO_TILES
  return;
#endif
  *reinterpret_cast<uint4*>(tile + chunk * 2048 + (s >> 3) * 128 + (s & 7) * 16) =
      make_uint4(lp_pack_bf16x2(x0, x1), lp_pack_bf16x2(x2, x3), lp_pack_bf16x2(x4, x5), lp_pack_bf16x2(x6, x7));
}
template <int N>
LP_DEVICE void lp_tile_row(unsigned char* tile, int chunk0, int s, const float (&x)[N]) {
#pragma unroll
  for (int c = 0; c < N / 8; ++c)
    lp_tile8(tile, chunk0 + c, s, x[8 * c], x[8 * c + 1], x[8 * c + 2], x[8 * c + 3], x[8 * c + 4], x[8 * c + 5], x[8 * c + 6], x[8 * c + 7]);
}
// A row that goes BOTH into a dW tile (bf16, round-to-nearest) and into the A operand (hi + lo): the tile's conversion
// is the operand's hi part (round-to-nearest instead of the truncation of lp_split2: the residual is then at most half a
// bf16 ulp), so a pair costs one conversion less than lp_tile_row + lp_stage_row.
template <int N, int LO = 16>
LP_DEVICE void lp_tile_stage_row(unsigned char* tile, int chunk0, int s, unsigned taddr_a, const float (&x)[N]) {
  unsigned hi[N / 2], lo[N / 2];
#pragma unroll
  for (int j = 0; j < N / 2; ++j) {
    hi[j] = lp_pack_bf16x2(x[2 * j], x[2 * j + 1]);
    const float2 r = lp_sub2(lp_f2(x[2 * j], x[2 * j + 1]), lp_f2(__uint_as_float(hi[j] << 16), __uint_as_float(hi[j] & 0xffff0000u)));
    lo[j] = lp_pack_bf16x2(r.x, r.y);
  }
#ifndef LP_ABL_NO_TILES
#pragma unroll
  for (int c = 0; c < N / 8; ++c)
    *reinterpret_cast<uint4*>(tile + (chunk0 + c) * 2048 + (s >> 3) * 128 + (s & 7) * 16) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
#endif
  lp_tmem_st<N / 2>(taddr_a, hi);
  lp_tmem_st<N / 2>(taddr_a + LO, lo);
}
// the same with max(x, 0) folded into the conversion
template <int N>
LP_DEVICE void lp_tile_row_relu(unsigned char* tile, int chunk0, int s, const float (&x)[N]) {
#ifdef LP_ABL_NO_TILES
  return;
#endif
#pragma unroll
  for (int c = 0; c < N / 8; ++c)
    *reinterpret_cast<uint4*>(tile + (chunk0 + c) * 2048 + (s >> 3) * 128 + (s & 7) * 16) =
        make_uint4(lp_pack_bf16x2_relu(x[8 * c], x[8 * c + 1]), lp_pack_bf16x2_relu(x[8 * c + 2], x[8 * c + 3]),
                   lp_pack_bf16x2_relu(x[8 * c + 4], x[8 * c + 5]), lp_pack_bf16x2_relu(x[8 * c + 6], x[8 * c + 7]));
}
template <int N>
LP_DEVICE unsigned lp_mask_pos(const float (&x)[N]) {
  unsigned m = 0;
#pragma unroll
  for (int j = 0; j < N; ++j) m |= (x[j] > 0.f ? 1u : 0u) << j;
  return m;
}

// ReLU gate read back from the bf16 operand tile the activation was written to (this thread's own row;
// an activation is > 0 iff its bf16 image is non-zero): x[j] survives iff feature j of the tile row != 0
template <int N>
LP_DEVICE void lp_gate_row(float (&x)[N], const unsigned char* tile, int chunk0, int s) {
#pragma unroll
  for (int c = 0; c < N / 8; ++c) {
    const uint4 w = *reinterpret_cast<const uint4*>(tile + (chunk0 + c) * 2048 + (s >> 3) * 128 + (s & 7) * 16);
    const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) lp_gate2(ww[e], x[8 * c + 2 * e], x[8 * c + 2 * e + 1]);
  }
}

// adjoint of lp_gather_regs: the owner thread scatters its row into the grid gradient
template <int C, int CW = C>
LP_DEVICE void lp_splat_regs(const LpGridSet& G, float* grad, int b, float x, float y, float z, const float (&d)[CW], int ch0 = 0) {
#ifdef LP_ABL_NO_MEM
  if (d[0] == 1.2345f) grad[0] = x;
  return;
#endif
  for (int gi = 0; gi < G.n; ++gi) {
    int off[8];
    float w[8];
    const int nt = lp_taps_i32(G.g[gi], C, b, x, y, z, off, w);
    if (nt == 0) continue;
#pragma unroll
    for (int tp = 0; tp < 8; ++tp) {
      if (tp < nt) {
        const bool on = w[tp] != 0.f;
#pragma unroll
        for (int k = 0; k < CW / 4; ++k)
          lp_red_add4_if(on, grad + off[tp] + ch0 + 4 * k, w[tp] * d[4 * k], w[tp] * d[4 * k + 1], w[tp] * d[4 * k + 2],
                         w[tp] * d[4 * k + 3]);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// Quad-transposed, footprint-merging scatter (the backward kernel's memory warps).
//
// A texel row is C floats = 4 x (C/4) contiguous floats.  With one thread per sample every `red.global.add.v4` of a
// warp touches 32 different rows, 16 bytes each: 32 partial sectors per instruction, and the L2 reduction rate
// (requests, not bytes) bounds the whole backward (profiles/ncu_r2b_bwd.md: ~1.1 reduced words / clock / SM).  Here
// the four lanes of a quad (four consecutive rays) first transpose their gradient rows, so that lane q holds channels
// [q*C/4, (q+1)*C/4) of all four samples; a tap is then ONE contiguous row per quad (full sectors, a quarter of the
// requests), and samples of the quad that share a bilinear footprint on a plane -- neighbouring pixels mostly do --
// are summed in registers first and reduced once (reference: grid_sample_util.py:40-99 issues one atomic per sample,
// tap and channel).  Warp-collective: all 32 lanes call it; `valid` says whether this lane's sample contributes.
// -------------------------------------------------------------------------------------------------------------------
template <int CW>
LP_DEVICE void lp_quad_transpose(float (&blk)[4][CW], int q) {
  // in: blk[k] = channel block k of MY sample; out: blk[j] = MY channel block (q) of the quad's sample j
#pragma unroll
  for (int k0 = 0; k0 < 4; k0 += 2) {  // round 1, partner lane ^ 1: the pair (k0, k0+1)
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      const float send = (q & 1) ? blk[k0][i] : blk[k0 + 1][i];
      const float recv = __shfl_xor_sync(LP_FULL_MASK, send, 1);
      if (q & 1) blk[k0][i] = recv; else blk[k0 + 1][i] = recv;
    }
  }
#pragma unroll
  for (int k0 = 0; k0 < 2; ++k0) {     // round 2, partner lane ^ 2: the pair (k0, k0+2)
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      const float send = (q & 2) ? blk[k0][i] : blk[k0 + 2][i];
      const float recv = __shfl_xor_sync(LP_FULL_MASK, send, 2);
      if (q & 2) blk[k0][i] = recv; else blk[k0 + 2][i] = recv;
    }
  }
}
template <int CW>
LP_DEVICE void lp_red_row(float* p, const float (&v)[CW]) {
#pragma unroll
  for (int k = 0; k < CW / 4; ++k) lp_red_add4(p + 4 * k, v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
// one grid's share of lp_splat_quad: this lane's taps (off, w; nt == 0: none) with footprint key `fkey`; NT = taps
// the unrolled passes cover (4: planes only), ntw = taps of this grid (warp-uniform)
template <int CW, int NT>
LP_DEVICE void lp_splat_quad_grid(float* grad, int q, int qbase, int nt, int fkey, int ntw, int (&off)[NT], float (&w)[NT],
                                  const float (&blk)[4][CW]) {
  if (!__any_sync(LP_FULL_MASK, nt != 0)) return;  // nobody in the warp touches this grid
  if (nt == 0) {
#pragma unroll
    for (int t = 0; t < NT; ++t) { off[t] = 0; w[t] = 0.f; }
  }
  const int key = nt != 0 ? fkey : -1;  // -1: no contribution
  int kA = max(key, __shfl_xor_sync(LP_FULL_MASK, key, 1));
  kA = max(kA, __shfl_xor_sync(LP_FULL_MASK, kA, 2));
  const bool inA = key >= 0 && key == kA;  // merged pass: every sample of the quad standing on footprint kA
  const unsigned mA = (__ballot_sync(LP_FULL_MASK, inA) >> qbase) & 15u;
  if (__any_sync(LP_FULL_MASK, mA != 0)) {
    const int owner = qbase + (mA ? __ffs((int)mA) - 1 : 0);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (t < ntw) {
        const int o = __shfl_sync(LP_FULL_MASK, off[t], owner);
        float v[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wj = __shfl_sync(LP_FULL_MASK, inA ? w[t] : 0.f, qbase + j);
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = fmaf(wj, blk[j][i], v[i]);
        }
        if (mA) lp_red_row<CW>(grad + o + q * CW, v);
      }
    }
  }
  // samples of the quad on another footprint: one pass each (rare for neighbouring pixels)
  const bool left = key >= 0 && !inA;
  const unsigned mL = __ballot_sync(LP_FULL_MASK, left);
  if (mL == 0) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!((mL >> j) & 0x11111111u)) continue;  // no quad has a leftover in position j
    const bool need = (mL >> (qbase + j)) & 1u;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (t < ntw) {
        const int o = __shfl_sync(LP_FULL_MASK, off[t], qbase + j);
        const float wj = __shfl_sync(LP_FULL_MASK, w[t], qbase + j);
        if (need) {
          float v[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = wj * blk[j][i];
          lp_red_row<CW>(grad + o + q * CW, v);
        }
      }
    }
  }
}
template <int C, bool TRI = true>
LP_DEVICE void lp_splat_quad(const LpGridSet& G, float* grad, int b, float x, float y, float z, bool valid, const float (&d)[C]) {
  constexpr int CW = C / 4;
  const int lane = threadIdx.x & 31, q = lane & 3, qbase = lane & ~3;
  float blk[4][CW];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < CW; ++i) blk[k][i] = valid ? d[k * CW + i] : 0.f;
  lp_quad_transpose<CW>(blk, q);
  if (TRI && G.tri) {  // triplane: axes once, one (run-time) loop over the three planes with four-tap passes
    const int W = G.g[0].W, Hh = G.g[0].H, Dd = G.g[1].D;
    const LpAxis ax = lp_axis_pre(x, W), ay = lp_axis_pre(y, Hh), az = lp_axis_pre(z, Dd);
#pragma unroll 1
    for (int p = 0; p < 3; ++p) {
      const LpAxis au = lp_axis_sel(p == 2, ay, ax), av = lp_axis_sel(p == 0, ay, az);
      const int U = p == 2 ? Hh : W, V = p == 0 ? Hh : Dd;
      const int nt = valid && au.ok && av.ok ? 4 : 0;
      int off[4];
      float w[4];
      lp_plane_taps_pre((int)G.g[p].base + b * U * V * C, U, C, au, av, off, w);
      lp_splat_quad_grid<CW, 4>(grad, q, qbase, nt, (b * (V + 3) + av.i0 + 2) * (U + 3) + au.i0 + 2, 4, off, w, blk);
    }
    return;
  }
  for (int gi = 0; gi < G.n; ++gi) {
    int off[8];
    float w[8];
    int fkey = -1;
    const int nt = valid ? lp_taps_i32(G.g[gi], C, b, x, y, z, off, w, &fkey) : 0;
    lp_splat_quad_grid<CW, 8>(grad, q, qbase, nt, fkey, G.g[gi].kind == LP_VOXEL ? 8 : 4, off, w, blk);
  }
}


}  // namespace lptc
