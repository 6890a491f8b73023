// TEST INFRASTRUCTURE ONLY -- a tiny host-side SIMT emulator for the CUDA-core (non-tcgen05) kernels of biapy_b200.
//
// The GPU-less build container cannot run a kernel, but index arithmetic, shuffle patterns and reduction layouts of the
// plain SIMT kernels can be checked on the CPU: tests/test_simt_emulation.py cuts the kernel *source text* out of
// biapy_b200/csrc/*.cu, pastes it behind this header and compiles it with g++.  Every CUDA thread of a block is a
// std::thread; __syncthreads() is a block barrier, __shfl_xor_sync() an exchange through a per-warp buffer between two warp
// barriers (so a divergent shuffle dead-locks here exactly like it would misbehave on the device), atomicAdd() takes a
// mutex.  Blocks run one after the other.  Nothing in the product links against this file.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3e { unsigned x = 1, y = 1, z = 1; };
static thread_local dim3e threadIdx, blockIdx, blockDim, gridDim;

// ---- 16-bit types -------------------------------------------------------------------------------------------------
struct __nv_bfloat16 { uint16_t x; };
static inline float __bfloat162float(__nv_bfloat16 v) {
  uint32_t u = (uint32_t)v.x << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return __nv_bfloat16{(uint16_t)0x7fff};
  u += 0x7fffu + ((u >> 16) & 1u);
  return __nv_bfloat16{(uint16_t)(u >> 16)};
}

// ---- block / warp machinery ---------------------------------------------------------------------------------------
struct BlockCtx {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<float> xchg;   // [warps][32]
  std::vector<double> dyn_smem_store;   // dynamic shared memory of the block (8-byte aligned)
  char* dyn_smem = nullptr;
};
static BlockCtx* g_ctx = nullptr;
static std::mutex g_atomic_mu;

static inline void __syncthreads() { g_ctx->block_bar->arrive_and_wait(); }

static inline float __shfl_xor_sync(unsigned, float v, int o) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  g_ctx->xchg[warp * 32 + lane] = v;
  g_ctx->warp_bar[warp]->arrive_and_wait();
  const float r = g_ctx->xchg[warp * 32 + (lane ^ o)];
  g_ctx->warp_bar[warp]->arrive_and_wait();
  return r;
}

static inline double __shfl_xor_sync(unsigned, double v, int o) {
  // two float-sized halves through the same exchange buffer would need care; use a second, double-typed buffer instead
  static std::vector<double> xd(64 * 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  xd[warp * 32 + lane] = v;
  g_ctx->warp_bar[warp]->arrive_and_wait();
  const double r = xd[warp * 32 + (lane ^ o)];
  g_ctx->warp_bar[warp]->arrive_and_wait();
  return r;
}

static inline float atomicAdd(float* p, float v) {
  std::lock_guard<std::mutex> l(g_atomic_mu);
  const float old = *p;
  *p = old + v;
  return old;
}

static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) {
  std::lock_guard<std::mutex> l(g_atomic_mu);
  const uint32_t old = *p;
  *p = old + v;
  return old;
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  std::lock_guard<std::mutex> l(g_atomic_mu);
  const unsigned long long old = *p;
  *p = old + v;
  return old;
}
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline double atomicAdd(double* p, double v) {
  std::lock_guard<std::mutex> l(g_atomic_mu);
  const double old = *p;
  *p = old + v;
  return old;
}
static inline float emu_expf(float x) { return std::exp(x); }
static inline float emu_fdividef(float a, float b) { return a / b; }
#define __expf(x) emu_expf(x)          // glibc declares a __expf of its own
#define __fdividef(a, b) emu_fdividef(a, b)
static inline uint32_t __float_as_uint(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}

// rounding-mode intrinsics of the stitch kernels: plain IEEE single operations (this file is compiled without FMA contraction)
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }

// launch(grid, block, [&]{ kernel(args...); }); emu_launch2 adds gridDim.y and dynamic shared memory
static inline void emu_launch2(unsigned grid_x, unsigned grid_y, unsigned block, size_t dyn_smem_bytes, const std::function<void()>& body);
static inline void emu_launch(unsigned grid, unsigned block, const std::function<void()>& body) { emu_launch2(grid, 1, block, 0, body); }
static inline void emu_launch2(unsigned grid, unsigned grid_y, unsigned block, size_t dyn_smem_bytes, const std::function<void()>& body) {
 for (unsigned by = 0; by < grid_y; ++by)
  for (unsigned b = 0; b < grid; ++b) {
    BlockCtx ctx;
    ctx.dyn_smem_store.assign(dyn_smem_bytes / 8 + 1, 0.0);
    ctx.dyn_smem = reinterpret_cast<char*>(ctx.dyn_smem_store.data());
    ctx.block_bar = std::make_unique<std::barrier<>>(block);
    const unsigned warps = (block + 31) / 32;
    for (unsigned w = 0; w < warps; ++w) {
      const unsigned lanes = (w + 1) * 32 <= block ? 32 : block - w * 32;
      ctx.warp_bar.push_back(std::make_unique<std::barrier<>>(lanes));
    }
    ctx.xchg.assign(warps * 32, 0.f);
    g_ctx = &ctx;
    std::vector<std::thread> ts;
    ts.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      ts.emplace_back([&, t] {
        threadIdx.x = t;
        blockIdx.x = b;
        blockIdx.y = by;
        blockDim.x = block;
        gridDim.x = grid;
        gridDim.y = grid_y;
        body();
      });
    for (auto& th : ts) th.join();
    g_ctx = nullptr;
  }
}

// ---- the element helpers of common.cuh / ops.cu (same definitions) ---------------------------------------------------
static inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static inline double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T> inline float to_f(T v);
template <> inline float to_f<float>(float v) { return v; }
template <> inline float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> inline T from_f(float v);
template <> inline float from_f<float>(float v) { return v; }
template <> inline __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

template <typename T, int VEC>
inline void load_vec(const T* p, float (&f)[VEC]) {
  Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(p);
  for (int i = 0; i < VEC; ++i) f[i] = to_f<T>(v.v[i]);
}
template <typename T, int VEC>
inline void store_vec(T* p, const float (&f)[VEC]) {
  Pack<T, VEC> v;
  for (int i = 0; i < VEC; ++i) v.v[i] = from_f<T>(f[i]);
  *reinterpret_cast<Pack<T, VEC>*>(p) = v;
}

struct PoolGeom {
  int n, d, h, w, c;
  int od, oh, ow;
  int pd, ph, pw;
};
constexpr int kSmallMax = 8;
