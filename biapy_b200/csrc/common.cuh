// Shared helpers for the biapy_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <math.h>

#include "../../include/biapy_b200.h"
#include <mutex>
#include <vector>

#define B200_EXPORT extern "C" __attribute__((visibility("default")))

namespace b200 {

void set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)                      \
  do {                                                  \
    if (!(cond)) {                                      \
      b200::set_error(__VA_ARGS__);                     \
      return B200_ERR_ARG;                              \
    }                                                   \
  } while (0)

#define B200_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      b200::set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return B200_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

#define B200_LAUNCH_CHECK() B200_CUDA(cudaGetLastError())

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------- element I/O
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// Vector of VEC elements of T moved with one 64/128-bit access.
template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

template <typename T> struct VecOf;                 // elements per 16 bytes
template <> struct VecOf<float> { static constexpr int n = 4; };
template <> struct VecOf<__nv_bfloat16> { static constexpr int n = 8; };
template <> struct VecOf<__half> { static constexpr int n = 8; };

__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case B200_ACT_RELU: return x > 0.f ? x : 0.f;
    case B200_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case B200_ACT_SILU: return __fdividef(x, 1.f + __expf(-x));
    case B200_ACT_LEAKY_RELU: return x > 0.f ? x : 0.01f * x;
    case B200_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    case B200_ACT_TANH: return tanhf(x);
    case B200_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case B200_ACT_SOFTPLUS: return x > 20.f ? x : log1pf(expf(x));
    default: return x;
  }
}

// derivative of the activation w.r.t. its input x (pre-activation value)
__device__ __forceinline__ float act_grad(int act, float x) {
  switch (act) {
    case B200_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case B200_ACT_ELU: return x > 0.f ? 1.f : expf(x);
    case B200_ACT_SILU: {
      float s = __fdividef(1.f, 1.f + __expf(-x));
      return s * (1.f + x * (1.f - s));
    }
    case B200_ACT_LEAKY_RELU: return x > 0.f ? 1.f : 0.01f;
    case B200_ACT_GELU: {
      float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
      float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
      return cdf + x * pdf;
    }
    case B200_ACT_TANH: {
      float t = tanhf(x);
      return 1.f - t * t;
    }
    case B200_ACT_SIGMOID: {
      float s = 1.f / (1.f + expf(-x));
      return s * (1.f - s);
    }
    case B200_ACT_SOFTPLUS: return 1.f / (1.f + expf(-x));
    default: return 1.f;
  }
}

// Compile-time activation (ACT >= 0) or the runtime switch (ACT == -1).  The HBM-bound row kernels are instantiated for
// the activations the U-Net family uses; a per-element jump table costs them more than the memory traffic does.
template <int ACT> __device__ __forceinline__ float act_fwd_t(int act, float x) { return act_fwd(ACT >= 0 ? ACT : act, x); }
template <int ACT> __device__ __forceinline__ float act_grad_t(int act, float x) { return act_grad(ACT >= 0 ? ACT : act, x); }

#define B200_DISPATCH_ACT(act, ACT, ...)                                                    \
  switch (act) {                                                                            \
    case B200_ACT_NONE: { constexpr int ACT = B200_ACT_NONE; __VA_ARGS__; break; }          \
    case B200_ACT_RELU: { constexpr int ACT = B200_ACT_RELU; __VA_ARGS__; break; }          \
    case B200_ACT_ELU: { constexpr int ACT = B200_ACT_ELU; __VA_ARGS__; break; }            \
    case B200_ACT_SILU: { constexpr int ACT = B200_ACT_SILU; __VA_ARGS__; break; }          \
    default: { constexpr int ACT = -1; __VA_ARGS__; break; }                                \
  }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Opt-in dynamic shared memory cap of a kernel, raised ONCE to the hardware maximum (227 KB per block minus the kernel's
// static shared memory).  The cap is a property of the function, not of a launch: a CUDA-graph kernel node replayed on its
// own (ncu's per-node profiling) sees whatever value the last eager launch left behind, so it must never be lowered.
template <typename K>
inline cudaError_t raise_dyn_smem_cap(K kern) {
  static std::mutex mu;
  static std::vector<const void*> done;
  std::lock_guard<std::mutex> lock(mu);
  for (const void* k : done)
    if (k == (const void*)kern) return cudaSuccess;
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kern);
  if (e != cudaSuccess) return e;
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
  if (e == cudaSuccess) done.push_back((const void*)kern);
  return e;
}

inline bool valid_dtype(int dt) { return dt == B200_F32 || dt == B200_BF16 || dt == B200_F16; }
inline size_t dtype_size(int dt) { return dt == B200_F32 ? 4 : 2; }

inline bool check_tensor(const b200_tensor* t, const char* name) {
  if (!t || !t->data) { set_error("%s: null tensor", name); return false; }
  if (!valid_dtype(t->dtype)) { set_error("%s: bad dtype %d", name, t->dtype); return false; }
  if (t->n <= 0 || t->d <= 0 || t->h <= 0 || t->w <= 0 || t->c <= 0 || t->ld < t->c) {
    set_error("%s: bad shape (%d,%d,%d,%d,%d) ld=%lld", name, t->n, t->d, t->h, t->w, t->c, (long long)t->ld);
    return false;
  }
  return true;
}
inline bool same_spatial(const b200_tensor* a, const b200_tensor* b) {
  return a->n == b->n && a->d == b->d && a->h == b->h && a->w == b->w;
}
inline int64_t voxels(const b200_tensor* t) { return (int64_t)t->n * t->d * t->h * t->w; }

// dispatch a templated functor on the runtime dtype
#define B200_DISPATCH_DTYPE(dt, T, ...)                                         \
  switch (dt) {                                                                 \
    case B200_F32: { using T = float; __VA_ARGS__; break; }                     \
    case B200_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }            \
    case B200_F16: { using T = __half; __VA_ARGS__; break; }                    \
    default: b200::set_error("bad dtype %d", (int)(dt)); return B200_ERR_ARG;   \
  }

#define B200_DISPATCH_DTYPE16(dt, T, ...)                                       \
  switch (dt) {                                                                 \
    case B200_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }            \
    case B200_F16: { using T = __half; __VA_ARGS__; break; }                    \
    default: b200::set_error("16-bit dtype expected, got %d", (int)(dt)); return B200_ERR_ARG; \
  }

}  // namespace b200
