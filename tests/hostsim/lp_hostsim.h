// TEST INFRASTRUCTURE -- SIMT emulation shim.
//
// Lets the CUDA sources under lightplane_b200/csrc/ compile with g++ (-DLP_HOSTSIM) so that the
// kernels' logic (indexing, MLP forward/backward, compositing, atomics, warp collectives) can be
// debugged against the oracle in the GPU-less build container.  Every CUDA thread of a block is a
// real OS thread; __syncthreads / warp collectives are barriers.  It is slow (tests use a few
// hundred rays), it is NOT a CPU fallback: the product library (liblightplane_b200.so) is never
// built with LP_HOSTSIM and lp_is_device_build() reports which one a .so is.
#pragma once

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __noinline__ __attribute__((noinline))
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "hostsim"; }

namespace lp_hostsim {

struct WarpCtx {
  std::barrier<>* bar;
  uint32_t slots_u[32];
  int pred[32];
  std::atomic<unsigned> gcnt[32] = {};  // rendezvous counters of sub-warp groups (keyed by the group's lowest lane)
};

struct BlockCtx {
  std::barrier<>* bar;
  unsigned char* smem;
  std::vector<WarpCtx>* warps;
  float* tmem;  // emulated tensor memory: [128 lanes][512 columns]
  std::mutex named_mu;
  std::unique_ptr<std::barrier<>> named[16];
  std::atomic<int> red[16][3];
};

struct ThreadCtx {
  uint3 tid, bid;
  dim3 bdim, gdim;
  BlockCtx* block;
  WarpCtx* warp;
  int lane;
};

inline thread_local ThreadCtx* g_ctx = nullptr;

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::vector<unsigned char> smem(smem_bytes + 64);
        std::barrier<> block_bar(nthreads);
        std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
        std::vector<WarpCtx> warps(nwarps);
        for (unsigned w = 0; w < nwarps; ++w) {
          unsigned cnt = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
          warp_bars.emplace_back(new std::barrier<>(cnt));
          warps[w].bar = warp_bars.back().get();
        }
        unsigned char* smem_aligned =
            reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 63) & ~uintptr_t(63));
        std::vector<float> tmem(128 * 512, 0.f);
        BlockCtx bctx;
        bctx.bar = &block_bar; bctx.smem = smem_aligned; bctx.warps = &warps; bctx.tmem = tmem.data();
        for (auto& r : bctx.red) { r[0] = 0; r[1] = 0; r[2] = 0; }
        std::vector<std::thread> threads;
        threads.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t) {
          threads.emplace_back([&, t]() {
            ThreadCtx ctx;
            ctx.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            ctx.bid = uint3{bx, by, bz};
            ctx.bdim = block;
            ctx.gdim = grid;
            ctx.block = &bctx;
            ctx.warp = &warps[t / 32];
            ctx.lane = t % 32;
            g_ctx = &ctx;
            body();
            g_ctx = nullptr;
          });
        }
        for (auto& th : threads) th.join();
      }
}

inline unsigned char* dyn_smem() { return g_ctx->block->smem; }

}  // namespace lp_hostsim

#define threadIdx (lp_hostsim::g_ctx->tid)
#define blockIdx (lp_hostsim::g_ctx->bid)
#define blockDim (lp_hostsim::g_ctx->bdim)
#define gridDim (lp_hostsim::g_ctx->gdim)

#define LP_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(lp_hostsim::dyn_smem())
#define LP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  lp_hostsim::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })

static inline void __syncthreads() { lp_hostsim::g_ctx->block->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { lp_hostsim::g_ctx->warp->bar->arrive_and_wait(); }

// Rendezvous of the lanes named by `mask` (a strict subset of the warp: sub-warp shuffles / ballots
// executed while other sub-warps of the same warp are elsewhere).  Every group of one warp keeps
// the same mask for the whole kernel, so a monotonic counter per lowest lane is a reusable barrier.
static inline void lp_hs_group_sync(unsigned mask) {
  auto* w = lp_hostsim::g_ctx->warp;
  if (mask == 0xffffffffu) { w->bar->arrive_and_wait(); return; }
  const unsigned n = (unsigned)__builtin_popcount(mask);
  std::atomic<unsigned>& c = w->gcnt[__builtin_ctz(mask)];
  const unsigned target = (c.fetch_add(1, std::memory_order_acq_rel) / n + 1) * n;
  while (c.load(std::memory_order_acquire) < target) std::this_thread::yield();
}
template <class T>
static inline T lp_hs_exchange(T v, int src_lane, unsigned mask = 0xffffffffu) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  auto* w = lp_hostsim::g_ctx->warp;
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  w->slots_u[lp_hostsim::g_ctx->lane] = bits;
  lp_hs_group_sync(mask);
  uint32_t got = w->slots_u[src_lane & 31];
  lp_hs_group_sync(mask);
  T out;
  std::memcpy(&out, &got, 4);
  return out;
}
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = lp_hostsim::g_ctx->lane;
  return lp_hs_exchange(v, (lane & ~(width - 1)) | (src & (width - 1)), mask);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
  return lp_hs_exchange(v, lp_hostsim::g_ctx->lane ^ m);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
  int l = lp_hostsim::g_ctx->lane + d;
  return lp_hs_exchange(v, l > 31 ? lp_hostsim::g_ctx->lane : l);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  auto* w = lp_hostsim::g_ctx->warp;
  w->pred[lp_hostsim::g_ctx->lane] = pred ? 1 : 0;
  lp_hs_group_sync(mask);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i)
    if ((mask >> i) & 1u) m |= (w->pred[i] ? 1u : 0u) << i;
  lp_hs_group_sync(mask);
  return m;
}
static inline int __reduce_min_sync(unsigned, int v) {
  int m = v;
  for (int d = 16; d > 0; d >>= 1) { int o = lp_hs_exchange(m, lp_hostsim::g_ctx->lane ^ d); m = o < m ? o : m; }
  return m;
}
static inline int __reduce_max_sync(unsigned, int v) {
  int m = v;
  for (int d = 16; d > 0; d >>= 1) { int o = lp_hs_exchange(m, lp_hostsim::g_ctx->lane ^ d); m = o > m ? o : m; }
  return m;
}
static inline int __ffs(int x) { return x ? __builtin_ffs(x) : 0; }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }

static inline float atomicAdd(float* p, float v) {
  return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed);
}
static inline int atomicAdd(int* p, int v) {
  return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_relaxed);
}

template <class T>
static inline T __ldg(const T* p) { return *p; }
// (glibc already declares __expf/__logf; kernels use expf/logf)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }

static inline float lp_hs_trunc_tf32(float x) {
  unsigned u; std::memcpy(&u, &x, 4); u &= 0xffffe000u; float y; std::memcpy(&y, &u, 4); return y;
}

// ---- emulation of the tcgen05 / mbarrier layer of lp_platform.cuh (protocol + layout logic) ----
// mbarrier: the 64-bit word counts completed phases (every barrier here expects one arrival).
// mbarrier emulation: low 32 bits = completed phases, bits 32..47 = arrivals of the current phase,
// bits 48..63 = expected arrivals per phase
static inline void lp_mbar_init(unsigned long long* bar, int count) { *bar = (unsigned long long)count << 48; }
static inline void lp_mbar_init_fence() {}
static inline void lp_mbar_wait(unsigned long long* bar, int parity) {
  unsigned long spins = 0;
  while ((int)(std::atomic_ref<unsigned long long>(*bar).load(std::memory_order_acquire) & 1ull) == parity) {
    std::this_thread::yield();
    if (++spins == (1ul << 22) && getenv("LP_HOSTSIM_WATCHDOG")) {  // debugging aid: report a stuck wait once
      const unsigned long long v = std::atomic_ref<unsigned long long>(*bar).load();
      fprintf(stderr, "[hostsim] thread %u block %u stuck on mbarrier at smem+%ld parity %d (phases %llu arrived %llu expect %llu)\n",
              lp_hostsim::g_ctx->tid.x, lp_hostsim::g_ctx->bid.x, (long)((unsigned char*)bar - lp_hostsim::g_ctx->block->smem), parity,
              v & 0xffffffffull, (v >> 32) & 0xffffull, v >> 48);
    }
  }
}
static inline void lp_hs_mbar_arrive(unsigned long long* bar) {
  std::atomic_ref<unsigned long long> a(*bar);
  unsigned long long o = a.load(std::memory_order_acquire), n;
  do {
    const unsigned long long expect = o >> 48, arrived = ((o >> 32) & 0xffffull) + 1;
    n = arrived == expect ? ((o & 0xffff0000ffffffffull) + 1) : (o + (1ull << 32));
  } while (!a.compare_exchange_weak(o, n, std::memory_order_acq_rel));
}
static inline void lp_mbar_arrive(unsigned long long* bar) { lp_hs_mbar_arrive(bar); }
static inline bool lp_elect_one() { return lp_hostsim::g_ctx->lane == 0; }
#define LP_SETMAXNREG_INC(n)
#define LP_SETMAXNREG_DEC(n)
static inline void lp_fence_async_smem() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void lp_tc_fence_before() {}
static inline void lp_tc_fence_after() {}
static inline void lp_tmem_alloc512(unsigned* slot) { if (lp_hostsim::g_ctx->lane == 0) *slot = 0; }
static inline void lp_tmem_dealloc512(unsigned) {}
static inline float lp_hs_bf16(const unsigned char* base, int byte_off) {
  unsigned short h; std::memcpy(&h, base + byte_off, 2);
  unsigned u = (unsigned)h << 16; float f; std::memcpy(&f, &u, 4); return f;
}
// MMAs execute synchronously (program order == issue order, like the in-order tensor pipe)
static inline void lp_tc_commit(unsigned long long* bar) {
  std::atomic_ref<unsigned long long> a(*bar);
  unsigned long long o = a.load(std::memory_order_acquire), n;
  do {
    const unsigned long long expect = o >> 48, arrived = ((o >> 32) & 0xffffull) + 1;
    n = arrived == expect ? ((o & 0xffff0000ffffffffull) + 1) : (o + (1ull << 32));
  } while (!a.compare_exchange_weak(o, n, std::memory_order_acq_rel));
}
// ---- TS MMAs / tcgen05.st / named barriers (see lp_platform.cuh); TMEM words hold raw 32-bit patterns ----
static inline unsigned lp_hs_tmem_word(int lane, int col) {
  unsigned u; std::memcpy(&u, &lp_hostsim::g_ctx->block->tmem[lane * 512 + col], 4); return u;
}
struct LpHsKDesc { const unsigned char* p; };
static inline const unsigned char* lp_tc_kdesc_lo(const void* smem_ptr) { return static_cast<const unsigned char*>(smem_ptr); }
typedef const unsigned char* lp_kdesc_t;
static inline lp_kdesc_t lp_tc_kadv(lp_kdesc_t p, int bytes) { return p + bytes; }
static inline void lp_tc_mma_ts(bool tf32, unsigned d_taddr, unsigned a_taddr, const unsigned char* b, int nstride, int n,
                                int accumulate) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  float* T = lp_hostsim::g_ctx->block->tmem;
  const int dcol = d_taddr & 0xffff, acol = a_taddr & 0xffff;
  const int K = tf32 ? 8 : 16;
  for (int m = 0; m < 128; ++m) {
    float a[16];
    for (int k = 0; k < K; ++k) {
      if (tf32) {
        a[k] = lp_hs_trunc_tf32(T[m * 512 + acol + k]);
      } else {
        const unsigned w = lp_hs_tmem_word(m, acol + k / 2);
        const unsigned u = (k & 1) ? (w & 0xffff0000u) : (w << 16);
        std::memcpy(&a[k], &u, 4);
      }
    }
    for (int j = 0; j < n; ++j) {
      float acc = accumulate ? T[m * 512 + dcol + j] : 0.f;
      for (int k = 0; k < K; ++k) {
        float bv;
        if (tf32) {
          std::memcpy(&bv, b + (j / 8) * nstride + (k / 4) * 128 + (j % 8) * 16 + (k % 4) * 4, 4);
          bv = lp_hs_trunc_tf32(bv);
        } else {
          bv = lp_hs_bf16(b, (j / 8) * nstride + (k / 8) * 128 + (j % 8) * 16 + (k % 8) * 2);
        }
        acc += a[k] * bv;
      }
      T[m * 512 + dcol + j] = acc;
    }
  }
}
static inline lp_kdesc_t lp_tc_kdesc_lo_t(const void* smem_ptr, int) { return static_cast<const unsigned char*>(smem_ptr); }
static inline void lp_tc_mma_ts_t(unsigned d_taddr, unsigned a_taddr, lp_kdesc_t b, int nstride, int n, int accumulate) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  float* T = lp_hostsim::g_ctx->block->tmem;
  const int dcol = d_taddr & 0xffff, acol = a_taddr & 0xffff;
  for (int m = 0; m < 128; ++m) {
    float a[16];
    for (int k = 0; k < 16; ++k) {
      const unsigned w = lp_hs_tmem_word(m, acol + k / 2);
      const unsigned u = (k & 1) ? (w & 0xffff0000u) : (w << 16);
      std::memcpy(&a[k], &u, 4);
    }
    for (int j = 0; j < n; ++j) {
      float acc = accumulate ? T[m * 512 + dcol + j] : 0.f;
      for (int k = 0; k < 16; ++k) acc += a[k] * lp_hs_bf16(b, (k / 8) * nstride + (j / 8) * 128 + (k % 8) * 16 + (j % 8) * 2);
      T[m * 512 + dcol + j] = acc;
    }
  }
}
static inline lp_kdesc_t lp_tc_mndesc_lo(const void* smem_ptr) { return static_cast<const unsigned char*>(smem_ptr); }
static inline void lp_tc_mma_ss_mn(unsigned d_taddr, lp_kdesc_t a, lp_kdesc_t b, int sbo, int n, int accumulate) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  float* T = lp_hostsim::g_ctx->block->tmem;
  const int dcol = d_taddr & 0xffff;
  auto off = [sbo](int mn, int k) { return (mn / 8) * sbo + (k / 8) * 128 + (k % 8) * 16 + (mn % 8) * 2; };
  for (int m = 0; m < 128; ++m)
    for (int j = 0; j < n; ++j) {
      float acc = accumulate ? T[m * 512 + dcol + j] : 0.f;
      for (int k = 0; k < 16; ++k) acc += lp_hs_bf16(a, off(m, k)) * lp_hs_bf16(b, off(j, k));
      T[m * 512 + dcol + j] = acc;
    }
}
static inline unsigned lp_taddr(unsigned tmem_base, int warp_in_group, int col) {
  return tmem_base + ((unsigned)(warp_in_group * 32) << 16) + (unsigned)col;
}
template <int NW>
static inline void lp_tmem_st(unsigned taddr, const unsigned (&v)[NW]) {
  float* T = lp_hostsim::g_ctx->block->tmem;
  const int row = (int)(taddr >> 16) + lp_hostsim::g_ctx->lane, col = taddr & 0xffff;
  for (int j = 0; j < NW; ++j) std::memcpy(&T[row * 512 + col + j], &v[j], 4);
}
template <int N>
static inline void lp_tmem_zero(unsigned taddr) {
  unsigned z[N];
  for (int j = 0; j < N; ++j) z[j] = 0u;
  lp_tmem_st<N>(taddr, z);
}
static inline void lp_tmem_wait_st() {}
static inline void lp_tmem_ld32u(unsigned taddr, float (&v)[32]) {
  const float* T = lp_hostsim::g_ctx->block->tmem;
  const int row = (int)(taddr >> 16) + lp_hostsim::g_ctx->lane, col = taddr & 0xffff;
  for (int j = 0; j < 32; ++j) v[j] = T[row * 512 + col + j];
}
template <int N>
static inline void lp_tmem_ld(unsigned taddr, float (&v)[N]) {
  const float* T = lp_hostsim::g_ctx->block->tmem;
  const int row = (int)(taddr >> 16) + lp_hostsim::g_ctx->lane, col = taddr & 0xffff;
  for (int j = 0; j < N; ++j) v[j] = T[row * 512 + col + j];
}
static inline void lp_bar_sync(int id, int nthreads) {
  auto* b = lp_hostsim::g_ctx->block;
  std::barrier<>* bar;
  {
    std::lock_guard<std::mutex> lk(b->named_mu);
    if (!b->named[id]) b->named[id].reset(new std::barrier<>(nthreads));
    bar = b->named[id].get();
  }
  if (getenv("LP_HOSTSIM_TRACE")) fprintf(stderr, "[hostsim] t%u bar.sync id %d n %d\n", lp_hostsim::g_ctx->tid.x, id, nthreads);
  bar->arrive_and_wait();
}
static inline bool lp_bar_any(int id, int nthreads, bool pred) {
  static thread_local unsigned gen[16] = {0};
  auto* b = lp_hostsim::g_ctx->block;
  const int slot = (int)(gen[id]++ % 3u);
  b->red[id][(slot + 1) % 3].store(0);  // next vote's slot: last read two votes ago, i.e. before the previous barrier
  if (pred) b->red[id][slot].fetch_or(1);
  lp_bar_sync(id, nthreads);
  return b->red[id][slot].load() != 0;
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  unsigned long long v = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
