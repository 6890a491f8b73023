// CUDA-core (FFMA, fp32 accumulate) convolution kernels: the exact-precision path (fp32 storage) and the
// fallback for shapes the tcgen05 kernels do not take (odd channel counts, first layer Cin=1..3, 1-channel
// heads, 5x5x5 "larger_io" kernels, 2D).  Same C ABI and packed-weight layout as the tensor-core path.
//   fprop : y[n,z,y,x,co] = b[co] + sum_{tap,ci} x[n,z+dz,y+dy,x+dx,ci] * w[co][tap][ci]     (+ residual)
//   dgrad : the same kernel with the flipped/transposed packing of the weights
//   wgrad : dw[co][tap][ci] += sum_vox dy[vox][co] * x[vox+tap][ci]
#include "common.cuh"
#include <stdlib.h>

namespace b200 {

struct ConvGeom {
  int n, d, h, w, cin, cout;
  int kd, kh, kw;
  int64_t ldx, ldy, ldr;
};

constexpr int kTW = 32;   // tile width (x), one lane per column
constexpr int kTH = 4;    // rows (y) per thread
constexpr int kTN = 32;   // output channels per block (4 warps x 8)

template <typename T>
__global__ void __launch_bounds__(128)
conv_fprop_simt_kernel(const T* __restrict__ x, const T* __restrict__ wp, const float* __restrict__ bias,
                       const T* __restrict__ res, T* __restrict__ y, ConvGeom g, int accumulate, int CK) {
  extern __shared__ float smem[];
  const int taps = g.kd * g.kh * g.kw;
  const int HW = kTW + g.kw - 1, HH = kTH + g.kh - 1;
  const int plane = g.kd * HH * HW;               // halo elements per input channel
  float* s_in = smem;                              // [CK][kd][HH][HW]
  float* s_w = smem + (size_t)CK * plane;          // [taps][CK][kTN]

  const int tiles_x = (g.w + kTW - 1) / kTW;
  const int x0 = (blockIdx.x % tiles_x) * kTW;
  const int y0 = (blockIdx.x / tiles_x) * kTH;
  const int z0 = blockIdx.y % g.d;
  const int n = blockIdx.y / g.d;
  const int co0 = blockIdx.z * kTN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pd = g.kd / 2, ph = g.kh / 2, pw = g.kw / 2;

  float acc[kTH][8];
#pragma unroll
  for (int r = 0; r < kTH; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;

  for (int ci0 = 0; ci0 < g.cin; ci0 += CK) {
    __syncthreads();
    // ---- input halo (channel fastest in global memory)
    for (int i = tid; i < plane * CK; i += blockDim.x) {
      int ck = i % CK;
      int p = i / CK;
      int xx = p % HW;
      int yy = (p / HW) % HH;
      int dz = p / (HW * HH);
      int gz = z0 + dz - pd, gy = y0 + yy - ph, gx = x0 + xx - pw, c = ci0 + ck;
      float v = 0.f;
      if (gz >= 0 && gz < g.d && gy >= 0 && gy < g.h && gx >= 0 && gx < g.w && c < g.cin)
        v = to_f<T>(x[((((int64_t)n * g.d + gz) * g.h + gy) * g.w + gx) * g.ldx + c]);
      s_in[(size_t)ck * plane + p] = v;
    }
    // ---- weights [tap][ck][n]
    for (int i = tid; i < taps * CK * kTN; i += blockDim.x) {
      int ck = i % CK;
      int t = (i / CK) % taps;
      int nn = i / (CK * taps);
      int co = co0 + nn, c = ci0 + ck;
      float v = 0.f;
      if (co < g.cout && c < g.cin) v = to_f<T>(wp[((int64_t)co * taps + t) * g.cin + c]);
      s_w[((size_t)t * CK + ck) * kTN + nn] = v;
    }
    __syncthreads();
    for (int dz = 0; dz < g.kd; ++dz)
      for (int dy = 0; dy < g.kh; ++dy)
        for (int dx = 0; dx < g.kw; ++dx) {
          const int t = (dz * g.kh + dy) * g.kw + dx;
          for (int ck = 0; ck < CK; ++ck) {
            const float4 w0 = *reinterpret_cast<const float4*>(&s_w[((size_t)t * CK + ck) * kTN + warp * 8]);
            const float4 w1 = *reinterpret_cast<const float4*>(&s_w[((size_t)t * CK + ck) * kTN + warp * 8 + 4]);
            const float* src = s_in + (size_t)ck * plane + ((size_t)dz * HH + dy) * HW + lane + dx;
#pragma unroll
            for (int r = 0; r < kTH; ++r) {
              const float a = src[r * HW];
              acc[r][0] = fmaf(a, w0.x, acc[r][0]);
              acc[r][1] = fmaf(a, w0.y, acc[r][1]);
              acc[r][2] = fmaf(a, w0.z, acc[r][2]);
              acc[r][3] = fmaf(a, w0.w, acc[r][3]);
              acc[r][4] = fmaf(a, w1.x, acc[r][4]);
              acc[r][5] = fmaf(a, w1.y, acc[r][5]);
              acc[r][6] = fmaf(a, w1.z, acc[r][6]);
              acc[r][7] = fmaf(a, w1.w, acc[r][7]);
            }
          }
        }
  }
  // ---- epilogue
  const int gx = x0 + lane;
  if (gx < g.w) {
#pragma unroll
    for (int r = 0; r < kTH; ++r) {
      const int gy = y0 + r;
      if (gy >= g.h) continue;
      const int64_t vox = (((int64_t)n * g.d + z0) * g.h + gy) * g.w + gx;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int co = co0 + warp * 8 + j;
        if (co >= g.cout) continue;
        float v = acc[r][j];
        if (bias) v += bias[co];
        if (res) v += to_f<T>(res[vox * g.ldr + co]);
        T* o = y + vox * g.ldy + co;
        if (accumulate) v += to_f<T>(*o);
        *o = from_f<T>(v);
      }
    }
  }
}

// --------------------------------------------------------------------------------------------------- wgrad
constexpr int kWgTile = 16;   // ci x co tile per block
constexpr int kWgVox = 32;    // voxels (along x) per step
constexpr int kWgTaps = 27;   // taps kept in registers per thread

template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                       float* __restrict__ dbias, ConvGeom g, int taps_per_chunk, int64_t units) {
  extern __shared__ float smem[];
  const int taps = g.kd * g.kh * g.kw;
  const int HW = kWgVox + g.kw - 1;
  const int t0 = blockIdx.z * taps_per_chunk;
  const int ntaps = min(taps_per_chunk, taps - t0);
  float* s_dy = smem;                                   // [kWgVox][16]
  float* s_x = smem + kWgVox * kWgTile;                  // [kd][kh][HW][16]
  __shared__ int s_off[kWgTaps];
  const int tid = threadIdx.x;
  const int ci_l = tid / kWgTile, co_l = tid % kWgTile;
  const int tiles_co = (g.cout + kWgTile - 1) / kWgTile;
  const int co0 = (blockIdx.y % tiles_co) * kWgTile;
  const int ci0 = (blockIdx.y / tiles_co) * kWgTile;
  const int pd = g.kd / 2, ph = g.kh / 2, pw = g.kw / 2;
  if (tid < kWgTaps) {
    int t = t0 + tid;
    int off = 0;
    if (tid < ntaps) {
      int dx = t % g.kw, dyy = (t / g.kw) % g.kh, dz = t / (g.kw * g.kh);
      off = ((dz * g.kh + dyy) * HW + dx) * kWgTile;
    }
    s_off[tid] = off;
  }
  float acc[kWgTaps];
#pragma unroll
  for (int t = 0; t < kWgTaps; ++t) acc[t] = 0.f;
  float bacc = 0.f;
  const int segs = (g.w + kWgVox - 1) / kWgVox;
  const int rows_x = g.kd * g.kh * HW;

  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    int seg = (int)(u % segs);
    int64_t r = u / segs;
    int yy = (int)(r % g.h); r /= g.h;
    int zz = (int)(r % g.d);
    int n = (int)(r / g.d);
    int x0 = seg * kWgVox;
    __syncthreads();
    for (int i = tid; i < kWgVox * kWgTile; i += blockDim.x) {
      int c = i % kWgTile, v = i / kWgTile;
      float val = 0.f;
      if (x0 + v < g.w && co0 + c < g.cout)
        val = to_f<T>(dy[((((int64_t)n * g.d + zz) * g.h + yy) * g.w + x0 + v) * g.ldy + co0 + c]);
      s_dy[i] = val;
    }
    for (int i = tid; i < rows_x * kWgTile; i += blockDim.x) {
      int c = i % kWgTile;
      int p = i / kWgTile;
      int xx = p % HW;
      int dyy = (p / HW) % g.kh;
      int dz = p / (HW * g.kh);
      int gz = zz + dz - pd, gy = yy + dyy - ph, gx = x0 + xx - pw;
      float val = 0.f;
      if (gz >= 0 && gz < g.d && gy >= 0 && gy < g.h && gx >= 0 && gx < g.w && ci0 + c < g.cin)
        val = to_f<T>(x[((((int64_t)n * g.d + gz) * g.h + gy) * g.w + gx) * g.ldx + ci0 + c]);
      s_x[i] = val;
    }
    __syncthreads();
    for (int v = 0; v < kWgVox; ++v) {
      const float d = s_dy[v * kWgTile + co_l];
      const float* xb = s_x + v * kWgTile + ci_l;
      bacc += d;
#pragma unroll
      for (int t = 0; t < kWgTaps; ++t)
        if (t < ntaps) acc[t] = fmaf(d, xb[s_off[t]], acc[t]);
    }
  }
  const int co = co0 + co_l, ci = ci0 + ci_l;
  if (co < g.cout && ci < g.cin) {
#pragma unroll
    for (int t = 0; t < kWgTaps; ++t)
      if (t < ntaps) atomicAdd(&dw[((int64_t)co * taps + t0 + t) * g.cin + ci], acc[t]);
  }
  if (dbias && ci0 == 0 && ci_l == 0 && blockIdx.z == 0 && co < g.cout) atomicAdd(&dbias[co], bacc);
}

// dbias[co] += sum_vox dy[vox][co]: block-level partial sums, one fp32 atomic per (block, channel)
template <typename T>
__global__ void bias_grad_kernel(const T* __restrict__ dy, int64_t ld, int c, int64_t nvox, float* __restrict__ dbias) {
  extern __shared__ float s_part[];   // [rows][c]
  const int rows = blockDim.x / c;
  const int row = threadIdx.x / c, ch = threadIdx.x % c;
  float acc = 0.f;
  if (row < rows)
    for (int64_t v = (int64_t)blockIdx.x * rows + row; v < nvox; v += (int64_t)gridDim.x * rows) acc += to_f<T>(dy[v * ld + ch]);
  if (row < rows) s_part[row * c + ch] = acc;
  __syncthreads();
  if (threadIdx.x < c) {
    float t = 0.f;
    for (int r = 0; r < rows; ++r) t += s_part[r * c + threadIdx.x];
    atomicAdd(&dbias[threadIdx.x], t);
  }
}

// vectorised form for 16-byte aligned channels-last rows: thread = (voxel row, 8-channel vector), four packed loads in flight
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 4) bias_grad_vec_kernel(const T* __restrict__ dy, int64_t ld, int c, int64_t nvox,
                                                              float* __restrict__ dbias, int cvn, int rows) {
  extern __shared__ float s_part[];   // [rows][c]
  const int cv = threadIdx.x % cvn, row = threadIdx.x / cvn;
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  if (row < rows) {
    const int64_t chunk = (nvox + gridDim.x - 1) / gridDim.x;
    const int64_t v0 = (int64_t)blockIdx.x * chunk;
    const int64_t v1 = v0 + chunk < nvox ? v0 + chunk : nvox;
    const T* base = dy + (int64_t)cv * VEC;
    constexpr int U = 4;
    for (int64_t v = v0 + row; v < v1; v += (int64_t)U * rows) {
      Pack<T, VEC> pk[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v + (int64_t)u * rows < v1) pk[u] = *reinterpret_cast<const Pack<T, VEC>*>(base + (v + (int64_t)u * rows) * ld);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v + (int64_t)u * rows < v1) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] += to_f<T>(pk[u].v[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) s_part[row * c + cv * VEC + i] = acc[i];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < rows; ++r) t += s_part[r * c + ch];
    atomicAdd(&dbias[ch], t);
  }
}

int conv_bias_grad(const b200_tensor* dy, float* dbias, cudaStream_t st) {
  B200_CHECK_ARG(dy->c <= 1024, "conv_bias_grad: too many channels");
  if (dy->dtype != B200_F32 && dy->c % 8 == 0 && dy->ld % 8 == 0 && dy->c / 8 <= 256 && ((uintptr_t)dy->data & 15) == 0) {
    const int cvn = dy->c / 8, rows = 256 / cvn;
    const int64_t nvox = voxels(dy);
    int64_t blocks = ceil_div(nvox, (int64_t)rows * 8);   // <= 2 rounds of 4 loads per thread on small tensors (latency-bound otherwise)
    if (blocks > (int64_t)sm_count() * 4) blocks = (int64_t)sm_count() * 4;
    const size_t smem = sizeof(float) * rows * dy->c;
    if (dy->dtype == B200_BF16)
      bias_grad_vec_kernel<__nv_bfloat16, 8><<<(unsigned)blocks, 256, smem, st>>>((const __nv_bfloat16*)dy->data, dy->ld, dy->c, nvox, dbias, cvn, rows);
    else
      bias_grad_vec_kernel<__half, 8><<<(unsigned)blocks, 256, smem, st>>>((const __half*)dy->data, dy->ld, dy->c, nvox, dbias, cvn, rows);
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  int threads = dy->c <= 256 ? 256 : 1024;
  threads = (threads / dy->c) * dy->c;
  int64_t nvox = voxels(dy);
  int blocks = (int)(ceil_div(nvox, threads / dy->c) < (int64_t)sm_count() * 8 ? ceil_div(nvox, threads / dy->c) : (int64_t)sm_count() * 8);
  B200_DISPATCH_DTYPE(dy->dtype, T, (bias_grad_kernel<T><<<blocks, threads, threads * sizeof(float), st>>>((const T*)dy->data, dy->ld, dy->c,
                                                                                                       nvox, dbias)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ p, int cout, int cin, int taps, int flip) {
  const int64_t total = (int64_t)cout * cin * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % taps);
    int ci = (int)((i / taps) % cin);
    int co = (int)(i / ((int64_t)taps * cin));
    float v = w[i];
    int64_t o = flip ? (((int64_t)ci * taps + (taps - 1 - t)) * cout + co) : (((int64_t)co * taps + t) * cin + ci);
    p[o] = from_f<T>(v);
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ p, float* __restrict__ dw, int cout, int cin, int taps,
                                    int accumulate) {
  const int64_t total = (int64_t)cout * cin * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % taps);
    int ci = (int)((i / taps) % cin);
    int co = (int)(i / ((int64_t)taps * cin));
    float v = p[((int64_t)co * taps + t) * cin + ci];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// ---------------------------------------------------------------------------------------- small pointwise convs
// k = 1 layers with a handful of channels on one side (segmentation heads 16->1, attention psi 8->1 and their
// gradients) are pure HBM streams: one thread per voxel, weights in shared memory, fp32 accumulation.
constexpr int kSmallMax = 8;

// y[vox][co] = b[co] + sum_ci x[vox][ci] * w[co][ci]   for cout <= 8 (any cin)
template <typename T>
__global__ void conv1x1_small_cout_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ wp, const float* __restrict__ bias,
                                          T* __restrict__ y, int64_t ldy, int cin, int cout, int64_t nvox, int accumulate) {
  extern __shared__ float s_w[];   // [cout][cin]
  for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) s_w[i] = to_f<T>(wp[i]);
  __syncthreads();
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * blockDim.x) {
    float acc[kSmallMax];
#pragma unroll
    for (int j = 0; j < kSmallMax; ++j) acc[j] = (bias && j < cout) ? bias[j] : 0.f;
    const T* xr = x + v * ldx;
    for (int ci = 0; ci < cin; ++ci) {
      float a = to_f<T>(xr[ci]);
#pragma unroll
      for (int j = 0; j < kSmallMax; ++j)
        if (j < cout) acc[j] = fmaf(a, s_w[j * cin + ci], acc[j]);
    }
    T* yr = y + v * ldy;
#pragma unroll
    for (int j = 0; j < kSmallMax; ++j)
      if (j < cout) yr[j] = from_f<T>(accumulate ? to_f<T>(yr[j]) + acc[j] : acc[j]);
  }
}

// y[vox][co] = b[co] + sum_{ci < cin <= 8} x[vox][ci] * w[co][ci]   for cin <= 8 (any cout)
template <typename T>
__global__ void conv1x1_small_cin_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ wp, const float* __restrict__ bias,
                                         T* __restrict__ y, int64_t ldy, int cin, int cout, int64_t nvox, int accumulate) {
  extern __shared__ float s_w[];   // [cout][cin]
  for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) s_w[i] = to_f<T>(wp[i]);
  __syncthreads();
  const int64_t total = nvox * cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = i / cout;
    int co = (int)(i % cout);
    float acc = bias ? bias[co] : 0.f;
    for (int ci = 0; ci < cin; ++ci) acc = fmaf(to_f<T>(x[v * ldx + ci]), s_w[co * cin + ci], acc);
    T* o = y + v * ldy + co;
    *o = from_f<T>(accumulate ? to_f<T>(*o) + acc : acc);
  }
}

// dw[co][ci] += sum_vox dy[vox][co] * x[vox][ci], dbias[co] += sum dy   for cin*cout <= 128 with min(cin,cout) <= 8
template <typename T>
__global__ void conv1x1_small_wgrad_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                                           float* __restrict__ dw, float* __restrict__ dbias, int cin, int cout, int64_t nvox) {
  // thread = (pair p = co*cin+ci) x voxel lane; 128 pairs max
  extern __shared__ float s_red[];
  const int pairs = cin * cout;
  const int lanes = blockDim.x / pairs;
  const int p = threadIdx.x % pairs, lane = threadIdx.x / pairs;
  const int co = p / cin, ci = p % cin;
  float acc = 0.f, bacc = 0.f;
  if (lane < lanes)
    for (int64_t v = (int64_t)blockIdx.x * lanes + lane; v < nvox; v += (int64_t)gridDim.x * lanes) {
      float d = to_f<T>(dy[v * lddy + co]);
      acc = fmaf(d, to_f<T>(x[v * ldx + ci]), acc);
      bacc += d;
    }
  s_red[threadIdx.x] = (lane < lanes) ? acc : 0.f;
  s_red[blockDim.x + threadIdx.x] = (lane < lanes) ? bacc : 0.f;
  __syncthreads();
  if (threadIdx.x < pairs) {
    float t = 0.f, tb = 0.f;
    for (int l = 0; l < lanes; ++l) {
      t += s_red[l * pairs + threadIdx.x];
      tb += s_red[blockDim.x + l * pairs + threadIdx.x];
    }
    atomicAdd(&dw[(int64_t)co * cin + ci], t);
    if (dbias && ci == 0) atomicAdd(&dbias[co], tb);
  }
}

// 16-bit, 16-byte-aligned forms: one thread per voxel, the wide side moves as 16-byte vectors (the scalar forms above
// issue one 2-byte access per channel and are LSU-bound at a fifth of the HBM rate)
// y[vox][co < cout <= 8] from cin = 8 * CV channels
template <typename T>
__global__ void __launch_bounds__(256) conv1x1_cout_vec_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ wp,
                                                               const float* __restrict__ bias, T* __restrict__ y, int64_t ldy, int cin,
                                                               int cout, int64_t nvox, int accumulate) {
  extern __shared__ float s_w[];   // [cout][cin]
  for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) s_w[i] = to_f<T>(wp[i]);
  __syncthreads();
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * blockDim.x) {
    float acc[kSmallMax];
#pragma unroll
    for (int j = 0; j < kSmallMax; ++j) acc[j] = (bias && j < cout) ? bias[j] : 0.f;
    const T* xr = x + v * ldx;
    for (int c0 = 0; c0 < cin; c0 += 8) {
      const Pack<T, 8> px = *reinterpret_cast<const Pack<T, 8>*>(xr + c0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = to_f<T>(px.v[i]);
#pragma unroll
        for (int j = 0; j < kSmallMax; ++j)
          if (j < cout) acc[j] = fmaf(a, s_w[j * cin + c0 + i], acc[j]);
      }
    }
    T* yr = y + v * ldy;
#pragma unroll
    for (int j = 0; j < kSmallMax; ++j)
      if (j < cout) yr[j] = from_f<T>(accumulate ? to_f<T>(yr[j]) + acc[j] : acc[j]);
  }
}

// y[vox][cout = 8 * CV] from cin <= 8 channels
template <typename T>
__global__ void __launch_bounds__(256) conv1x1_cin_vec_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ wp,
                                                              const float* __restrict__ bias, T* __restrict__ y, int64_t ldy, int cin,
                                                              int cout, int64_t nvox, int accumulate) {
  extern __shared__ float s_w[];   // [cout][cin]
  for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) s_w[i] = to_f<T>(wp[i]);
  __syncthreads();
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * blockDim.x) {
    float a[kSmallMax];
#pragma unroll
    for (int i = 0; i < kSmallMax; ++i) a[i] = i < cin ? to_f<T>(x[v * ldx + i]) : 0.f;
    T* yr = y + v * ldy;
    for (int c0 = 0; c0 < cout; c0 += 8) {
      Pack<T, 8> old;
      if (accumulate) old = *reinterpret_cast<const Pack<T, 8>*>(yr + c0);
      Pack<T, 8> out;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float acc = bias ? bias[c0 + j] : 0.f;
#pragma unroll
        for (int i = 0; i < kSmallMax; ++i)
          if (i < cin) acc = fmaf(a[i], s_w[(c0 + j) * cin + i], acc);
        if (accumulate) acc += to_f<T>(old.v[j]);
        out.v[j] = from_f<T>(acc);
      }
      *reinterpret_cast<Pack<T, 8>*>(yr + c0) = out;
    }
  }
}

// dw[co][ci], dbias[co] for cout <= 2 and cin = 8 * CV <= 32 (segmentation heads): per-thread fp32 partials, block reduce
template <typename T>
__global__ void __launch_bounds__(256) conv1x1_wgrad_head_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy,
                                                                 int64_t lddy, float* __restrict__ dw, float* __restrict__ dbias,
                                                                 int cin, int cout, int64_t nvox) {
  constexpr int kMaxCin = 32, kMaxCout = 2;
  __shared__ float s_red[8][kMaxCout * (kMaxCin + 1)];
  float acc[kMaxCout][kMaxCin], bacc[kMaxCout];
#pragma unroll
  for (int j = 0; j < kMaxCout; ++j) {
    bacc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxCin; ++i) acc[j][i] = 0.f;
  }
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * blockDim.x) {
    float d[kMaxCout];
#pragma unroll
    for (int j = 0; j < kMaxCout; ++j) d[j] = j < cout ? to_f<T>(dy[v * lddy + j]) : 0.f;
#pragma unroll
    for (int c0 = 0; c0 < kMaxCin; c0 += 8)
      if (c0 < cin) {
        const Pack<T, 8> px = *reinterpret_cast<const Pack<T, 8>*>(x + v * ldx + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a = to_f<T>(px.v[i]);
#pragma unroll
          for (int j = 0; j < kMaxCout; ++j) acc[j][c0 + i] = fmaf(d[j], a, acc[j][c0 + i]);
        }
      }
#pragma unroll
    for (int j = 0; j < kMaxCout; ++j) bacc[j] += d[j];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < kMaxCout; ++j) {
#pragma unroll
    for (int i = 0; i < kMaxCin; ++i) {
      const float t = warp_sum(acc[j][i]);
      if (lane == 0) s_red[warp][j * (kMaxCin + 1) + i] = t;
    }
    const float tb = warp_sum(bacc[j]);
    if (lane == 0) s_red[warp][j * (kMaxCin + 1) + kMaxCin] = tb;
  }
  __syncthreads();
  for (int it = threadIdx.x; it < kMaxCout * (kMaxCin + 1); it += blockDim.x) {
    const int j = it / (kMaxCin + 1), i = it % (kMaxCin + 1);
    if (j >= cout) continue;
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w][it];
    if (i < cin) atomicAdd(&dw[(int64_t)j * cin + i], t);
    else if (i == kMaxCin && dbias) atomicAdd(&dbias[j], t);
  }
}

// dw[co][ci], dbias[co] for cin <= 4 and cout = 8 or 16 (image-fed pointwise shortcut): thread per voxel, dY row as 16-byte
// vectors, fp32 partials per thread, warp shuffle + shared-memory block reduce, one atomic per (block, entry)
template <typename T>
__global__ void __launch_bounds__(256) conv1x1_wgrad_image_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy,
                                                                  int64_t lddy, float* __restrict__ dw, float* __restrict__ dbias,
                                                                  int cin, int cout, int64_t nvox) {
  constexpr int kCi = 4, kCo = 16;
  __shared__ float s_red[8][kCo * (kCi + 1)];
  float acc[kCi][kCo], bacc[kCo];
#pragma unroll
  for (int j = 0; j < kCo; ++j) {
    bacc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < kCi; ++i) acc[i][j] = 0.f;
  }
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (int64_t)gridDim.x * blockDim.x) {
    float a[kCi];
#pragma unroll
    for (int i = 0; i < kCi; ++i) a[i] = i < cin ? to_f<T>(x[v * ldx + i]) : 0.f;
#pragma unroll
    for (int c0 = 0; c0 < kCo; c0 += 8)
      if (c0 < cout) {
        const Pack<T, 8> pd = *reinterpret_cast<const Pack<T, 8>*>(dy + v * lddy + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = to_f<T>(pd.v[j]);
          bacc[c0 + j] += d;
#pragma unroll
          for (int i = 0; i < kCi; ++i) acc[i][c0 + j] = fmaf(d, a[i], acc[i][c0 + j]);
        }
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < kCo; ++j) {
#pragma unroll
    for (int i = 0; i < kCi; ++i) {
      const float t = warp_sum(acc[i][j]);
      if (lane == 0) s_red[warp][j * (kCi + 1) + i] = t;
    }
    const float tb = warp_sum(bacc[j]);
    if (lane == 0) s_red[warp][j * (kCi + 1) + kCi] = tb;
  }
  __syncthreads();
  for (int it = threadIdx.x; it < kCo * (kCi + 1); it += blockDim.x) {
    const int j = it / (kCi + 1), i = it % (kCi + 1);
    if (j >= cout) continue;
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w][it];
    if (i < cin) atomicAdd(&dw[(int64_t)j * cin + i], t);
    else if (i == kCi && dbias) atomicAdd(&dbias[j], t);
  }
}

// ---- coalesced forms (B200_PW_COALESCED=0 falls back to the thread-per-voxel kernels above) ---------------------------------
// thread = (voxel, 16-byte channel vector) of the WIDE tensor, so a warp instruction moves 512 contiguous bytes; the
// thread-per-voxel forms touch every 32-byte sector with two half-filled requests (stores) or two instructions (loads)
// and ran at 1.6-2.2 TB/s (0.13-0.18 ms per launch at 128^3 x 16 ch x batch 4 against a 0.045 ms HBM floor).
// CVN = vectors per voxel row, a power of two (the grid stride is a multiple of it, so a thread keeps its channel vector).

// y[vox][co < cout <= JMAX] from cin = 8 * CVN channels; the CVN partial dot products of a voxel meet by shuffle
template <typename T, int CVN, int JMAX>
__global__ void __launch_bounds__(256) conv1x1_cout_cv_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ wp,
                                                              const float* __restrict__ bias, T* __restrict__ y, int64_t ldy,
                                                              int cout, int64_t nvox, int accumulate) {
  constexpr int cin = 8 * CVN;
  const int cv = threadIdx.x & (CVN - 1);
  float w[JMAX][8];
#pragma unroll
  for (int j = 0; j < JMAX; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) w[j][i] = j < cout ? to_f<T>(wp[j * cin + cv * 8 + i]) : 0.f;
  const int64_t total = nvox * CVN, stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 2;   // two voxel vectors in flight per thread
  // warp-uniform trip count: the shuffles below need every lane
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < total; i0 += U * stride) {
    int64_t v[U];
    bool live[U];
    Pack<T, 8> px[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride + (threadIdx.x & 31);
      live[u] = i < total;
      v[u] = i / CVN;
      if (live[u]) px[u] = *reinterpret_cast<const Pack<T, 8>*>(x + v[u] * ldx + cv * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float acc[JMAX];
#pragma unroll
      for (int j = 0; j < JMAX; ++j) acc[j] = 0.f;
      if (live[u]) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float a = to_f<T>(px[u].v[k]);
#pragma unroll
          for (int j = 0; j < JMAX; ++j) acc[j] = fmaf(a, w[j][k], acc[j]);
        }
      }
#pragma unroll
      for (int o = CVN / 2; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < JMAX; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      if (live[u] && cv == 0) {
        T* yr = y + v[u] * ldy;
#pragma unroll
        for (int j = 0; j < JMAX; ++j)
          if (j < cout) {
            const float r = acc[j] + (bias ? bias[j] : 0.f);
            yr[j] = from_f<T>(accumulate ? to_f<T>(yr[j]) + r : r);
          }
      }
    }
  }
}

// y[vox][cout = 8 * cvn] from cin <= CMAX channels; cvn (a power of two) vectors per voxel, weights of the thread's vector
// in registers
template <typename T, int CMAX>
__global__ void __launch_bounds__(256) conv1x1_cin_cv_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ wp,
                                                             const float* __restrict__ bias, T* __restrict__ y, int64_t ldy, int cin,
                                                             int cvn_log2, int64_t nvox, int accumulate) {
  const int cvn = 1 << cvn_log2;
  const int cv = threadIdx.x & (cvn - 1);
  float w[8][CMAX], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    b[j] = bias ? bias[cv * 8 + j] : 0.f;
#pragma unroll
    for (int i = 0; i < CMAX; ++i) w[j][i] = i < cin ? to_f<T>(wp[(cv * 8 + j) * cin + i]) : 0.f;
  }
  const int64_t total = nvox << cvn_log2, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t v = i >> cvn_log2;
    float a[CMAX];
#pragma unroll
    for (int k = 0; k < CMAX; ++k) a[k] = k < cin ? to_f<T>(x[v * ldx + k]) : 0.f;
    T* yr = y + v * ldy + cv * 8;
    Pack<T, 8> old;
    if (accumulate) old = *reinterpret_cast<const Pack<T, 8>*>(yr);
    Pack<T, 8> out;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = b[j];
#pragma unroll
      for (int k = 0; k < CMAX; ++k) acc = fmaf(a[k], w[j][k], acc);
      if (accumulate) acc += to_f<T>(old.v[j]);
      out.v[j] = from_f<T>(acc);
    }
    *reinterpret_cast<Pack<T, 8>*>(yr) = out;
  }
}

// dw[co][ci], dbias[co] for cout <= MAXCO (2: segmentation heads; 8: the F_int = 8 layers of the attention gates) and
// cin = 8 * CVN <= 32
template <typename T, int CVN, int MAXCO = 2>
__global__ void __launch_bounds__(256) conv1x1_wgrad_head_cv_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy,
                                                                    int64_t lddy, float* __restrict__ dw, float* __restrict__ dbias,
                                                                    int cout, int64_t nvox) {
  constexpr int kMaxCin = 32, kMaxCout = MAXCO, cin = 8 * CVN;
  __shared__ float s_red[8][kMaxCout * (kMaxCin + 1)];
  const int cv = threadIdx.x & (CVN - 1);
  float acc[kMaxCout][8], bacc[kMaxCout];
#pragma unroll
  for (int j = 0; j < kMaxCout; ++j) {
    bacc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
  }
  const int64_t total = nvox * CVN, stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = MAXCO > 2 ? 2 : 4;   // voxel vectors in flight per thread (loads first, then the FMAs)
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
    Pack<T, 8> px[U];
    float d[U][kMaxCout];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      const bool live = i < total;
      const int64_t v = live ? i / CVN : 0;
      px[u] = *reinterpret_cast<const Pack<T, 8>*>(x + v * ldx + cv * 8);
#pragma unroll
      for (int j = 0; j < kMaxCout; ++j) d[u][j] = (live && j < cout) ? to_f<T>(dy[v * lddy + j]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float a = to_f<T>(px[u].v[k]);
#pragma unroll
        for (int j = 0; j < kMaxCout; ++j) acc[j][k] = fmaf(d[u][j], a, acc[j][k]);
      }
      if (cv == 0) {
#pragma unroll
        for (int j = 0; j < kMaxCout; ++j) bacc[j] += d[u][j];
      }
    }
  }
  // lanes with the same channel vector (lane % CVN) meet: xor offsets 16 .. CVN
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int it = threadIdx.x; it < 8 * kMaxCout * (kMaxCin + 1); it += blockDim.x) (&s_red[0][0])[it] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kMaxCout; ++j) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = acc[j][k];
#pragma unroll
      for (int o = 16; o >= CVN; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane < CVN) s_red[warp][j * (kMaxCin + 1) + lane * 8 + k] = t;
    }
    float tb = bacc[j];
#pragma unroll
    for (int o = 16; o >= CVN; o >>= 1) tb += __shfl_xor_sync(0xffffffffu, tb, o);
    if (lane == 0) s_red[warp][j * (kMaxCin + 1) + kMaxCin] = tb;
  }
  __syncthreads();
  for (int it = threadIdx.x; it < kMaxCout * (kMaxCin + 1); it += blockDim.x) {
    const int j = it / (kMaxCin + 1), k = it % (kMaxCin + 1);
    if (j >= cout) continue;
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w][it];
    if (k < cin) atomicAdd(&dw[(int64_t)j * cin + k], t);
    else if (k == kMaxCin && dbias) atomicAdd(&dbias[j], t);
  }
}

// dw[co][ci], dbias[co] for cin <= 4 and cout = 8 * CVN in {8, 16} (image-fed pointwise shortcut)
template <typename T, int CVN>
__global__ void __launch_bounds__(256) conv1x1_wgrad_image_cv_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy,
                                                                     int64_t lddy, float* __restrict__ dw, float* __restrict__ dbias,
                                                                     int cin, int64_t nvox) {
  constexpr int kCi = 4, kCo = 16, cout = 8 * CVN;
  __shared__ float s_red[8][kCo * (kCi + 1)];
  const int cv = threadIdx.x & (CVN - 1);
  float acc[kCi][8], bacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bacc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < kCi; ++k) acc[k][j] = 0.f;
  }
  const int64_t total = nvox * CVN, stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;   // four voxel vectors in flight per thread
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
    Pack<T, 8> pd[U];
    float a[U][kCi];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      live[u] = i < total;
      const int64_t v = live[u] ? i / CVN : 0;
      pd[u] = *reinterpret_cast<const Pack<T, 8>*>(dy + v * lddy + cv * 8);
#pragma unroll
      for (int k = 0; k < kCi; ++k) a[u][k] = k < cin ? to_f<T>(x[v * ldx + k]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = live[u] ? to_f<T>(pd[u].v[j]) : 0.f;
        bacc[j] += d;
#pragma unroll
        for (int k = 0; k < kCi; ++k) acc[k][j] = fmaf(d, a[u][k], acc[k][j]);
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int k = 0; k < kCi; ++k) {
      float t = acc[k][j];
#pragma unroll
      for (int o = 16; o >= CVN; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane < CVN) s_red[warp][(lane * 8 + j) * (kCi + 1) + k] = t;
    }
    float tb = bacc[j];
#pragma unroll
    for (int o = 16; o >= CVN; o >>= 1) tb += __shfl_xor_sync(0xffffffffu, tb, o);
    if (lane < CVN) s_red[warp][(lane * 8 + j) * (kCi + 1) + kCi] = tb;
  }
  __syncthreads();
  for (int it = threadIdx.x; it < cout * (kCi + 1); it += blockDim.x) {
    const int j = it / (kCi + 1), k = it % (kCi + 1);
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w][it];
    if (k < cin) atomicAdd(&dw[(int64_t)j * cin + k], t);
    else if (k == kCi && dbias) atomicAdd(&dbias[j], t);
  }
}

static inline bool pw_coalesced_enabled() {
  static const bool on = !(getenv("B200_PW_COALESCED") && atoi(getenv("B200_PW_COALESCED")) == 0);
  return on;
}
static inline unsigned pw_blocks(int64_t threads) {
  int64_t blocks = ceil_div(threads, 256);
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

template <typename T>
static void launch_cout_cv(const b200_tensor* x, const void* w, const float* bias, const b200_tensor* y, int64_t nvox, int accumulate,
                           cudaStream_t st) {
  const unsigned nb = pw_blocks(nvox * (x->c / 8));
  const T* xp = (const T*)x->data;
  const T* wp = (const T*)w;
  T* yp = (T*)y->data;
  const bool narrow = y->c <= 2;
  switch (x->c / 8) {
    case 1:
      if (narrow) conv1x1_cout_cv_kernel<T, 1, 2><<<nb, 256, 0, st>>>(xp, x->ld, wp, bias, yp, y->ld, y->c, nvox, accumulate);
      else conv1x1_cout_cv_kernel<T, 1, kSmallMax><<<nb, 256, 0, st>>>(xp, x->ld, wp, bias, yp, y->ld, y->c, nvox, accumulate);
      break;
    case 2:
      if (narrow) conv1x1_cout_cv_kernel<T, 2, 2><<<nb, 256, 0, st>>>(xp, x->ld, wp, bias, yp, y->ld, y->c, nvox, accumulate);
      else conv1x1_cout_cv_kernel<T, 2, kSmallMax><<<nb, 256, 0, st>>>(xp, x->ld, wp, bias, yp, y->ld, y->c, nvox, accumulate);
      break;
    default:
      if (narrow) conv1x1_cout_cv_kernel<T, 4, 2><<<nb, 256, 0, st>>>(xp, x->ld, wp, bias, yp, y->ld, y->c, nvox, accumulate);
      else conv1x1_cout_cv_kernel<T, 4, kSmallMax><<<nb, 256, 0, st>>>(xp, x->ld, wp, bias, yp, y->ld, y->c, nvox, accumulate);
      break;
  }
}

template <typename T>
static void launch_cin_cv(const b200_tensor* x, const void* w, const float* bias, const b200_tensor* y, int64_t nvox, int accumulate,
                          cudaStream_t st) {
  const int cvn = y->c / 8;
  int lg = 0;
  while ((1 << lg) < cvn) ++lg;
  const unsigned nb = pw_blocks(nvox * cvn);
  const T* xp = (const T*)x->data;
  const T* wp = (const T*)w;
  T* yp = (T*)y->data;
  if (x->c <= 2) conv1x1_cin_cv_kernel<T, 2><<<nb, 256, 0, st>>>(xp, x->ld, wp, bias, yp, y->ld, x->c, lg, nvox, accumulate);
  else if (x->c <= 4) conv1x1_cin_cv_kernel<T, 4><<<nb, 256, 0, st>>>(xp, x->ld, wp, bias, yp, y->ld, x->c, lg, nvox, accumulate);
  else conv1x1_cin_cv_kernel<T, 8><<<nb, 256, 0, st>>>(xp, x->ld, wp, bias, yp, y->ld, x->c, lg, nvox, accumulate);
}

template <typename T>
static void launch_wgrad_head_cv(const b200_tensor* x, const b200_tensor* dy, float* dw, float* dbias, int64_t nvox, cudaStream_t st) {
  const int cvn = x->c / 8;
  int64_t blocks = ceil_div(nvox * cvn, 256 * 8);
  if (blocks > (int64_t)sm_count() * 6) blocks = (int64_t)sm_count() * 6;
  if (blocks < 1) blocks = 1;
  const T* xp = (const T*)x->data;
  const T* gp = (const T*)dy->data;
  if (dy->c > 2) {
    if (cvn == 1) conv1x1_wgrad_head_cv_kernel<T, 1, 8><<<(unsigned)blocks, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dw, dbias, dy->c, nvox);
    else if (cvn == 2) conv1x1_wgrad_head_cv_kernel<T, 2, 8><<<(unsigned)blocks, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dw, dbias, dy->c, nvox);
    else conv1x1_wgrad_head_cv_kernel<T, 4, 8><<<(unsigned)blocks, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dw, dbias, dy->c, nvox);
    return;
  }
  if (cvn == 1) conv1x1_wgrad_head_cv_kernel<T, 1><<<(unsigned)blocks, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dw, dbias, dy->c, nvox);
  else if (cvn == 2) conv1x1_wgrad_head_cv_kernel<T, 2><<<(unsigned)blocks, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dw, dbias, dy->c, nvox);
  else conv1x1_wgrad_head_cv_kernel<T, 4><<<(unsigned)blocks, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dw, dbias, dy->c, nvox);
}

template <typename T>
static void launch_wgrad_image_cv(const b200_tensor* x, const b200_tensor* dy, float* dw, float* dbias, int64_t nvox, cudaStream_t st) {
  const int cvn = dy->c / 8;
  int64_t blocks = ceil_div(nvox * cvn, 256 * 8);
  if (blocks > (int64_t)sm_count() * 6) blocks = (int64_t)sm_count() * 6;
  if (blocks < 1) blocks = 1;
  const T* xp = (const T*)x->data;
  const T* gp = (const T*)dy->data;
  if (cvn == 1) conv1x1_wgrad_image_cv_kernel<T, 1><<<(unsigned)blocks, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dw, dbias, x->c, nvox);
  else conv1x1_wgrad_image_cv_kernel<T, 2><<<(unsigned)blocks, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dw, dbias, x->c, nvox);
}

static bool vec16(const b200_tensor* t) { return t->dtype != B200_F32 && t->c % 8 == 0 && t->ld % 8 == 0 && ((uintptr_t)t->data & 15) == 0; }

static bool small_pointwise(const b200_tensor* x, const b200_tensor* y, int kd, int kh, int kw) {
  return kd == 1 && kh == 1 && kw == 1 && (x->c <= kSmallMax || y->c <= kSmallMax) && x->c * y->c <= 128;
}

int conv_fprop_simt(const b200_tensor* x, const void* w, const float* bias, const b200_tensor* res, const b200_tensor* y,
                    int kd, int kh, int kw, int accumulate, cudaStream_t st) {
  if (!res && small_pointwise(x, y, kd, kh, kw)) {
    const int64_t nvox = voxels(x);
    const size_t smem = sizeof(float) * x->c * y->c;
    B200_DISPATCH_DTYPE(x->dtype, T, {
      if (y->c <= kSmallMax) {
        int64_t blocks = ceil_div(nvox, 256);
        if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
        if (pw_coalesced_enabled() && vec16(x) && (x->c == 8 || x->c == 16 || x->c == 32))
          launch_cout_cv<T>(x, w, bias, y, nvox, accumulate, st);
        else if (vec16(x))
          conv1x1_cout_vec_kernel<T><<<(unsigned)blocks, 256, smem, st>>>((const T*)x->data, x->ld, (const T*)w, bias, (T*)y->data,
                                                                        y->ld, x->c, y->c, nvox, accumulate);
        else
          conv1x1_small_cout_kernel<T><<<(unsigned)blocks, 256, smem, st>>>((const T*)x->data, x->ld, (const T*)w, bias, (T*)y->data,
                                                                          y->ld, x->c, y->c, nvox, accumulate);
      } else if (pw_coalesced_enabled() && vec16(y) && x->c <= 8 && ((y->c / 8) & (y->c / 8 - 1)) == 0) {
        launch_cin_cv<T>(x, w, bias, y, nvox, accumulate, st);
      } else if (vec16(y)) {
        int64_t blocks = ceil_div(nvox, 256);
        if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
        conv1x1_cin_vec_kernel<T><<<(unsigned)blocks, 256, smem, st>>>((const T*)x->data, x->ld, (const T*)w, bias, (T*)y->data,
                                                                     y->ld, x->c, y->c, nvox, accumulate);
      } else {
        int64_t blocks = ceil_div(nvox * y->c, 256);
        if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
        conv1x1_small_cin_kernel<T><<<(unsigned)blocks, 256, smem, st>>>((const T*)x->data, x->ld, (const T*)w, bias, (T*)y->data,
                                                                       y->ld, x->c, y->c, nvox, accumulate);
      }
    });
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  ConvGeom g{x->n, x->d, x->h, x->w, x->c, y->c, kd, kh, kw, x->ld, y->ld, res ? res->ld : 0};
  const int taps = kd * kh * kw;
  const size_t per_ck = ((size_t)kd * (kTH + kh - 1) * (kTW + kw - 1) + (size_t)taps * kTN) * sizeof(float);
  int CK = (int)(96 * 1024 / per_ck);
  if (CK > 8) CK = 8;
  if (CK > x->c) CK = x->c;
  B200_CHECK_ARG(CK >= 1, "conv_fprop(simt): kernel %dx%dx%d too large", kd, kh, kw);
  const size_t smem = per_ck * CK;
  dim3 grid((unsigned)(ceil_div(x->w, kTW) * ceil_div(x->h, kTH)), (unsigned)(x->d * x->n), (unsigned)ceil_div(y->c, kTN));
  B200_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "conv_fprop(simt): grid too large");
  B200_DISPATCH_DTYPE(x->dtype, T, {
    auto kern = conv_fprop_simt_kernel<T>;
    B200_CUDA(raise_dyn_smem_cap(kern));  // function-level cap: keep it at the maximum (graph nodes replayed alone)
    kern<<<grid, 128, smem, st>>>((const T*)x->data, (const T*)w, bias, res ? (const T*)res->data : nullptr, (T*)y->data, g,
                                  accumulate, CK);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int conv_wgrad_simt(const b200_tensor* x, const b200_tensor* dy, float* dw, float* dbias, int kd, int kh, int kw,
                    cudaStream_t st) {
  const bool head8 = pw_coalesced_enabled() && dy->c <= 8 && (x->c == 8 || x->c == 16 || x->c == 32) && x->dtype != B200_F32;
  if (small_pointwise(x, dy, kd, kh, kw) && (dy->c <= 2 || head8) && x->c <= 32 && vec16(x)) {
    const int64_t nvox = voxels(x);
    int64_t blocks = ceil_div(nvox, 256 * 8);
    if (blocks > (int64_t)sm_count() * 4) blocks = (int64_t)sm_count() * 4;
    if (blocks < 1) blocks = 1;
    if (pw_coalesced_enabled() && (x->c == 8 || x->c == 16 || x->c == 32)) {
      B200_DISPATCH_DTYPE(x->dtype, T, (launch_wgrad_head_cv<T>(x, dy, dw, dbias, nvox, st)));
    } else {
      B200_DISPATCH_DTYPE(x->dtype, T, (conv1x1_wgrad_head_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(
                                           (const T*)x->data, x->ld, (const T*)dy->data, dy->ld, dw, dbias, x->c, dy->c, nvox)));
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  if (small_pointwise(x, dy, kd, kh, kw) && x->c <= 4 && (dy->c == 8 || dy->c == 16) && vec16(dy) && x->dtype != B200_F32) {
    const int64_t nvox = voxels(x);
    int64_t blocks = ceil_div(nvox, 256 * 8);
    if (blocks > (int64_t)sm_count() * 4) blocks = (int64_t)sm_count() * 4;
    if (blocks < 1) blocks = 1;
    if (pw_coalesced_enabled()) {
      B200_DISPATCH_DTYPE(x->dtype, T, (launch_wgrad_image_cv<T>(x, dy, dw, dbias, nvox, st)));
    } else {
      B200_DISPATCH_DTYPE(x->dtype, T, (conv1x1_wgrad_image_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(
                                           (const T*)x->data, x->ld, (const T*)dy->data, dy->ld, dw, dbias, x->c, dy->c, nvox)));
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  if (small_pointwise(x, dy, kd, kh, kw)) {
    const int64_t nvox = voxels(x);
    const int pairs = x->c * dy->c;
    const int threads = (256 / pairs) * pairs >= pairs ? (256 / pairs) * pairs : pairs;
    int64_t blocks = ceil_div(nvox, (int64_t)(threads / pairs) * 16);
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    if (blocks < 1) blocks = 1;
    B200_DISPATCH_DTYPE(x->dtype, T, (conv1x1_small_wgrad_kernel<T><<<(unsigned)blocks, threads, 2 * threads * sizeof(float), st>>>(
                                         (const T*)x->data, x->ld, (const T*)dy->data, dy->ld, dw, dbias, x->c, dy->c, nvox)));
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  ConvGeom g{x->n, x->d, x->h, x->w, x->c, dy->c, kd, kh, kw, x->ld, dy->ld, 0};
  const int taps = kd * kh * kw;
  int tpc = taps <= kWgTaps ? taps : kh * kw;
  B200_CHECK_ARG(tpc <= kWgTaps, "conv_wgrad(simt): kernel plane %dx%d too large", kh, kw);
  const int chunks = (int)ceil_div(taps, tpc);
  const size_t smem = ((size_t)kWgVox * kWgTile + (size_t)kd * kh * (kWgVox + kw - 1) * kWgTile) * sizeof(float);
  const int64_t units = (int64_t)x->n * x->d * x->h * ceil_div(x->w, kWgVox);
  const int tiles = (int)(ceil_div(x->c, kWgTile) * ceil_div(dy->c, kWgTile));
  int64_t bx = ceil_div((int64_t)sm_count() * 4, (int64_t)tiles * chunks);
  if (bx > units) bx = units;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)tiles, (unsigned)chunks);
  B200_CHECK_ARG(grid.y <= 65535, "conv_wgrad(simt): too many channel tiles");
  B200_DISPATCH_DTYPE(x->dtype, T, {
    auto kern = conv_wgrad_simt_kernel<T>;
    B200_CUDA(raise_dyn_smem_cap(kern));  // function-level cap: keep it at the maximum (graph nodes replayed alone)
    kern<<<grid, 256, smem, st>>>((const T*)x->data, (const T*)dy->data, dw, dbias, g, tpc, units);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// implemented in conv_umma.cu
int conv_fprop_umma(const b200_tensor* x, const void* w, const float* bias, const b200_tensor* res, const b200_tensor* y,
                    int kd, int kh, int kw, int accumulate, cudaStream_t st);
bool conv_fprop_umma_supported(const b200_tensor* x, const b200_tensor* res, const b200_tensor* y, int kd, int kh, int kw);
int conv_wgrad_umma(const b200_tensor* x, const b200_tensor* dy, float* dw, float* dbias, int kd, int kh, int kw,
                    cudaStream_t st);
bool conv_wgrad_umma_supported(const b200_tensor* x, const b200_tensor* dy, int kd, int kh, int kw);
bool conv_fprop_xfold_supported(const b200_tensor* x, const b200_tensor* y, int kd, int kh, int kw);
int conv_fprop_xfold(const b200_tensor* x, const void* w, const float* bias, const b200_tensor* y, int kd, int kh, int kw,
                     int accumulate, cudaStream_t st);
int pack_weight_xfold(const float* w, void* packed, int dtype, int cout, int cin, int kd, int kh, int kw, int flip,
                      cudaStream_t st);

}  // namespace b200

using namespace b200;

B200_EXPORT int b200_pack_conv_weight(const float* w, void* packed, int32_t dtype, int32_t cout, int32_t cin, int32_t kd,
                                      int32_t kh, int32_t kw, int32_t flip_transpose, void* stream) {
  B200_CHECK_ARG(w && packed && cout > 0 && cin > 0 && kd > 0 && kh > 0 && kw > 0, "pack_conv_weight: bad args");
  const int taps = kd * kh * kw;
  int64_t total = (int64_t)cout * cin * taps;
  unsigned blocks = (unsigned)(ceil_div(total, 256) < 4096 ? ceil_div(total, 256) : 4096);
  B200_DISPATCH_DTYPE(dtype, T, (pack_weight_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (T*)packed, cout, cin, taps,
                                                                                              flip_transpose)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_unpack_conv_wgrad(const float* dw_packed, float* dw, int32_t cout, int32_t cin, int32_t taps,
                                       int32_t accumulate, void* stream) {
  B200_CHECK_ARG(dw_packed && dw && cout > 0 && cin > 0 && taps > 0, "unpack_conv_wgrad: bad args");
  int64_t total = (int64_t)cout * cin * taps;
  unsigned blocks = (unsigned)(ceil_div(total, 256) < 4096 ? ceil_div(total, 256) : 4096);
  unpack_wgrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dw_packed, dw, cout, cin, taps, accumulate);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static int check_conv(const b200_tensor* x, const b200_tensor* y, int kd, int kh, int kw, const char* who) {
  B200_CHECK_ARG(check_tensor(x, who) && check_tensor(y, who), "%s", b200_last_error());
  B200_CHECK_ARG(same_spatial(x, y) && x->dtype == y->dtype, "%s: x/y spatial shape or dtype mismatch", who);
  B200_CHECK_ARG(kd > 0 && kh > 0 && kw > 0 && (kd & 1) && (kh & 1) && (kw & 1), "%s: kernel sizes must be odd ('same' padding)", who);
  return B200_OK;
}

B200_EXPORT int b200_conv_fprop(const b200_tensor* x, const void* w_packed, const float* bias, const b200_tensor* residual,
                                const b200_tensor* y, int32_t kd, int32_t kh, int32_t kw, int32_t accumulate, int32_t impl,
                                void* stream) {
  int st = check_conv(x, y, kd, kh, kw, "conv_fprop");
  if (st) return st;
  B200_CHECK_ARG(w_packed != nullptr, "conv_fprop: null weights");
  if (residual)
    B200_CHECK_ARG(check_tensor(residual, "conv_fprop.residual") && same_spatial(residual, y) && residual->c == y->c &&
                       residual->dtype == y->dtype, "conv_fprop: residual mismatch");
  cudaStream_t s = (cudaStream_t)stream;
  if (impl == B200_IMPL_XFOLD) {   // weights must have been packed with b200_pack_conv_weight_xfold
    B200_CHECK_ARG(residual == nullptr && conv_fprop_xfold_supported(x, y, kd, kh, kw),
                   "conv_fprop: operands not supported by the x-folded kernel (query b200_conv_impl_query)");
    return conv_fprop_xfold(x, w_packed, bias, y, kd, kh, kw, accumulate, s);
  }
  bool umma_ok = conv_fprop_umma_supported(x, residual, y, kd, kh, kw);
  if (impl == B200_IMPL_UMMA && !umma_ok) {
    set_error("conv_fprop: shape not supported by the tcgen05 kernel (cin=%d cout=%d k=%dx%dx%d dtype=%d)", x->c, y->c, kd, kh,
              kw, x->dtype);
    return B200_ERR_UNSUPPORTED;
  }
  if (impl == B200_IMPL_UMMA || (impl == B200_IMPL_AUTO && umma_ok))
    return conv_fprop_umma(x, w_packed, bias, residual, y, kd, kh, kw, accumulate, s);
  return conv_fprop_simt(x, w_packed, bias, residual, y, kd, kh, kw, accumulate, s);
}

B200_EXPORT int b200_conv_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dw_packed, float* dbias, int32_t kd,
                                int32_t kh, int32_t kw, int32_t impl, void* stream) {
  int st = check_conv(x, dy, kd, kh, kw, "conv_wgrad");
  if (st) return st;
  B200_CHECK_ARG(dw_packed != nullptr, "conv_wgrad: null output");
  cudaStream_t s = (cudaStream_t)stream;
  bool umma_ok = conv_wgrad_umma_supported(x, dy, kd, kh, kw);
  if (impl == B200_IMPL_UMMA && !umma_ok) {
    set_error("conv_wgrad: shape not supported by the tcgen05 kernel");
    return B200_ERR_UNSUPPORTED;
  }
  if (impl == B200_IMPL_UMMA || (impl == B200_IMPL_AUTO && umma_ok)) return conv_wgrad_umma(x, dy, dw_packed, dbias, kd, kh, kw, s);
  return conv_wgrad_simt(x, dy, dw_packed, dbias, kd, kh, kw, s);
}

B200_EXPORT int b200_conv_impl_query(const b200_tensor* x, const b200_tensor* y, int32_t kd, int32_t kh, int32_t kw,
                                     int32_t wgrad) {
  if (!x || !y) return B200_IMPL_SIMT;
  if (!wgrad && conv_fprop_xfold_supported(x, y, kd, kh, kw)) return B200_IMPL_XFOLD;
  bool ok = wgrad ? conv_wgrad_umma_supported(x, y, kd, kh, kw) : conv_fprop_umma_supported(x, nullptr, y, kd, kh, kw);
  return ok ? B200_IMPL_UMMA : B200_IMPL_SIMT;
}

B200_EXPORT int b200_pack_conv_weight_xfold(const float* w, void* packed, int32_t dtype, int32_t cout, int32_t cin, int32_t kd,
                                            int32_t kh, int32_t kw, int32_t flip_transpose, void* stream) {
  B200_CHECK_ARG(w && packed && cout > 0 && cin > 0 && kd > 0 && kh > 0 && (kw == 1 || kw == 3), "pack_conv_weight_xfold: bad args");
  return pack_weight_xfold(w, packed, dtype, cout, cin, kd, kh, kw, flip_transpose, (cudaStream_t)stream);
}
