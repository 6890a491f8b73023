// Transposed convolution with kernel == stride (nn.ConvTranspose{2,3}d(k=s, stride=s), biapy/models/blocks.py:603,
// 1607): every input voxel produces an independent s_d x s_h x s_w block of output voxels, so
//   fprop : Y[vox, (tap,co)] = X[vox, ci] . W[ci, (co,tap)]        -- one GEMM with a scatter epilogue
//   dgrad : dX[vox, ci]      = dY[vox, (tap,co)] . W^T             -- one GEMM with a gather prologue
//   wgrad : dW[ci, co, tap]  = sum_vox X[vox, ci] * dY[up(vox,tap), co]
// CUDA-core fp32-accumulate kernels; weights are read in the PyTorch layout (Cin, Cout, sd, sh, sw) fp32.
#include "common.cuh"

namespace b200 {

struct ConvTGeom {
  int n, d, h, w, cin, cout;   // input spatial dims
  int sd, sh, sw;
  int64_t ldx, ldy;
};

__device__ __forceinline__ int64_t up_voxel(const ConvTGeom& g, int64_t vox, int tap) {
  int x = (int)(vox % g.w); int64_t t = vox / g.w;
  int y = (int)(t % g.h); t /= g.h;
  int z = (int)(t % g.d);
  int n = (int)(t / g.d);
  int c = tap % g.sw, b = (tap / g.sw) % g.sh, a = tap / (g.sw * g.sh);
  return ((((int64_t)n * g.d * g.sd + (int64_t)z * g.sd + a) * (g.h * g.sh) + (int64_t)y * g.sh + b) * (g.w * g.sw)) +
         (int64_t)x * g.sw + c;
}

constexpr int kCK = 16;

// block: 128 threads; 32 voxels (lanes) x 32 columns j=(tap,co) (4 warps x 8)
template <typename T>
__global__ void __launch_bounds__(128)
convT_fprop_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   T* __restrict__ y, ConvTGeom g, int64_t nvox) {
  __shared__ float s_x[kCK][33];
  __shared__ float s_w[kCK][32];
  const int taps = g.sd * g.sh * g.sw;
  const int ncols = taps * g.cout;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t v0 = (int64_t)blockIdx.x * 32;
  const int j0 = blockIdx.y * 32;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int ci0 = 0; ci0 < g.cin; ci0 += kCK) {
    __syncthreads();
    for (int i = tid; i < 32 * kCK; i += 128) {
      int ck = i % kCK, v = i / kCK;
      float val = 0.f;
      if (v0 + v < nvox && ci0 + ck < g.cin) val = to_f<T>(x[(v0 + v) * g.ldx + ci0 + ck]);
      s_x[ck][v] = val;
    }
    for (int i = tid; i < 32 * kCK; i += 128) {
      int jj = i % 32, ck = i / 32;
      int j = j0 + jj;
      float val = 0.f;
      if (j < ncols && ci0 + ck < g.cin) {
        int tap = j / g.cout, co = j % g.cout;
        val = w[((int64_t)(ci0 + ck) * g.cout + co) * taps + tap];
      }
      s_w[ck][jj] = val;
    }
    __syncthreads();
#pragma unroll
    for (int ck = 0; ck < kCK; ++ck) {
      float a = s_x[ck][lane];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(a, s_w[ck][warp * 8 + j], acc[j]);
    }
  }
  const int64_t vox = v0 + lane;
  if (vox < nvox) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int jc = j0 + warp * 8 + j;
      if (jc >= ncols) continue;
      int tap = jc / g.cout, co = jc % g.cout;
      float v = acc[j] + (bias ? bias[co] : 0.f);
      y[up_voxel(g, vox, tap) * g.ldy + co] = from_f<T>(v);
    }
  }
}

// block: 128 threads; 32 voxels x 32 input channels; K = (tap,co)
template <typename T>
__global__ void __launch_bounds__(128)
convT_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx, ConvTGeom g, int64_t nvox,
                   int accumulate) {
  __shared__ float s_a[kCK][33];
  __shared__ float s_w[kCK][32];
  const int taps = g.sd * g.sh * g.sw;
  const int K = taps * g.cout;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t v0 = (int64_t)blockIdx.x * 32;
  const int ci0 = blockIdx.y * 32;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kCK) {
    __syncthreads();
    for (int i = tid; i < 32 * kCK; i += 128) {
      int kk = i % kCK, v = i / kCK;
      int k = k0 + kk;
      float val = 0.f;
      if (v0 + v < nvox && k < K) {
        int tap = k / g.cout, co = k % g.cout;
        val = to_f<T>(dy[up_voxel(g, v0 + v, tap) * g.ldy + co]);
      }
      s_a[kk][v] = val;
    }
    for (int i = tid; i < 32 * kCK; i += 128) {
      int cc = i % 32, kk = i / 32;
      int k = k0 + kk, ci = ci0 + cc;
      float val = 0.f;
      if (k < K && ci < g.cin) {
        int tap = k / g.cout, co = k % g.cout;
        val = w[((int64_t)ci * g.cout + co) * taps + tap];
      }
      s_w[kk][cc] = val;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kCK; ++kk) {
      float a = s_a[kk][lane];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(a, s_w[kk][warp * 8 + j], acc[j]);
    }
  }
  const int64_t vox = v0 + lane;
  if (vox < nvox) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int ci = ci0 + warp * 8 + j;
      if (ci >= g.cin) continue;
      T* o = dx + vox * g.ldx + ci;
      float v = acc[j];
      if (accumulate) v += to_f<T>(*o);
      *o = from_f<T>(v);
    }
  }
}

constexpr int kMaxTTaps = 8;

// block: 256 threads = 16 ci x 16 co; grid.x = voxel splits, grid.y = channel tiles
template <typename T>
__global__ void __launch_bounds__(256)
convT_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, float* __restrict__ dbias,
                   ConvTGeom g, int64_t nvox) {
  __shared__ float s_x[32][16];
  __shared__ float s_dy[kMaxTTaps][32][16];
  const int taps = g.sd * g.sh * g.sw;
  const int tid = threadIdx.x;
  const int ci_l = tid / 16, co_l = tid % 16;
  const int tiles_co = (g.cout + 15) / 16;
  const int co0 = (blockIdx.y % tiles_co) * 16, ci0 = (blockIdx.y / tiles_co) * 16;
  float acc[kMaxTTaps];
#pragma unroll
  for (int t = 0; t < kMaxTTaps; ++t) acc[t] = 0.f;
  float bacc = 0.f;
  const int64_t units = (nvox + 31) / 32;
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    const int64_t v0 = u * 32;
    __syncthreads();
    for (int i = tid; i < 32 * 16; i += 256) {
      int c = i % 16, v = i / 16;
      float val = 0.f;
      if (v0 + v < nvox && ci0 + c < g.cin) val = to_f<T>(x[(v0 + v) * g.ldx + ci0 + c]);
      s_x[v][c] = val;
    }
    for (int i = tid; i < taps * 32 * 16; i += 256) {
      int c = i % 16, v = (i / 16) % 32, t = i / (16 * 32);
      float val = 0.f;
      if (v0 + v < nvox && co0 + c < g.cout) val = to_f<T>(dy[up_voxel(g, v0 + v, t) * g.ldy + co0 + c]);
      s_dy[t][v][c] = val;
    }
    __syncthreads();
    for (int v = 0; v < 32; ++v) {
      float a = s_x[v][ci_l];
#pragma unroll
      for (int t = 0; t < kMaxTTaps; ++t)
        if (t < taps) {
          float d = s_dy[t][v][co_l];
          acc[t] = fmaf(a, d, acc[t]);
          bacc += d;
        }
    }
  }
  const int ci = ci0 + ci_l, co = co0 + co_l;
  if (ci < g.cin && co < g.cout) {
#pragma unroll
    for (int t = 0; t < kMaxTTaps; ++t)
      if (t < taps) atomicAdd(&dw[((int64_t)ci * g.cout + co) * taps + t], acc[t]);
  }
  if (dbias && ci0 == 0 && ci_l == 0 && co < g.cout) atomicAdd(&dbias[co], bacc);
}

static int convT_check(const b200_tensor* x, const b200_tensor* y, int sd, int sh, int sw, const char* who) {
  B200_CHECK_ARG(check_tensor(x, who) && check_tensor(y, who), "%s", b200_last_error());
  B200_CHECK_ARG(sd > 0 && sh > 0 && sw > 0, "%s: bad stride", who);
  B200_CHECK_ARG(x->dtype == y->dtype && y->n == x->n && y->d == x->d * sd && y->h == x->h * sh && y->w == x->w * sw,
                 "%s: output shape must be input*stride", who);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

B200_EXPORT int b200_convT_fprop(const b200_tensor* x, const float* w, const float* bias, const b200_tensor* y, int32_t sd,
                                 int32_t sh, int32_t sw, void* stream) {
  int st = convT_check(x, y, sd, sh, sw, "convT_fprop");
  if (st) return st;
  B200_CHECK_ARG(w != nullptr, "convT_fprop: null weights");
  ConvTGeom g{x->n, x->d, x->h, x->w, x->c, y->c, sd, sh, sw, x->ld, y->ld};
  int64_t nvox = voxels(x);
  dim3 grid((unsigned)ceil_div(nvox, 32), (unsigned)ceil_div((int64_t)sd * sh * sw * y->c, 32));
  B200_DISPATCH_DTYPE(x->dtype, T, (convT_fprop_kernel<T><<<grid, 128, 0, (cudaStream_t)stream>>>((const T*)x->data, w, bias,
                                                                                              (T*)y->data, g, nvox)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_convT_dgrad(const b200_tensor* dy, const float* w, const b200_tensor* dx, int32_t sd, int32_t sh,
                                 int32_t sw, int32_t accumulate, void* stream) {
  int st = convT_check(dx, dy, sd, sh, sw, "convT_dgrad");
  if (st) return st;
  B200_CHECK_ARG(w != nullptr, "convT_dgrad: null weights");
  ConvTGeom g{dx->n, dx->d, dx->h, dx->w, dx->c, dy->c, sd, sh, sw, dx->ld, dy->ld};
  int64_t nvox = voxels(dx);
  dim3 grid((unsigned)ceil_div(nvox, 32), (unsigned)ceil_div(dx->c, 32));
  B200_DISPATCH_DTYPE(dx->dtype, T, (convT_dgrad_kernel<T><<<grid, 128, 0, (cudaStream_t)stream>>>((const T*)dy->data, w,
                                                                                               (T*)dx->data, g, nvox, accumulate)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_convT_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dw, float* dbias, int32_t sd, int32_t sh,
                                 int32_t sw, void* stream) {
  int st = convT_check(x, dy, sd, sh, sw, "convT_wgrad");
  if (st) return st;
  B200_CHECK_ARG(dw != nullptr, "convT_wgrad: null output");
  B200_CHECK_ARG(sd * sh * sw <= kMaxTTaps, "convT_wgrad: stride volume %d > %d not supported", sd * sh * sw, kMaxTTaps);
  ConvTGeom g{x->n, x->d, x->h, x->w, x->c, dy->c, sd, sh, sw, x->ld, dy->ld};
  int64_t nvox = voxels(x);
  int tiles = (int)(ceil_div(x->c, 16) * ceil_div(dy->c, 16));
  int64_t bx = ceil_div((int64_t)sm_count() * 4, tiles);
  int64_t units = ceil_div(nvox, 32);
  if (bx > units) bx = units;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)tiles);
  B200_DISPATCH_DTYPE(x->dtype, T, (convT_wgrad_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x->data,
                                                                                              (const T*)dy->data, dw, dbias, g, nvox)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}
