// HBM-bound kernels of the U-Net path: normalisation statistics / apply / backward, activations, max-pool,
// element-wise glue, losses and the optimiser step.  All tensors are channels-last with a voxel pitch `ld`.
// Reference call sites are listed next to each entry point in include/biapy_b200.h.
#include "common.cuh"
#include <stdlib.h>

#include <type_traits>

namespace b200 {

// View used inside kernels
template <typename T>
struct View {
  T* p;
  int64_t ld;
  int c;
  int64_t vox;      // n*d*h*w
  int64_t spatial;  // d*h*w
};
template <typename T>
static View<T> view(const b200_tensor* t) {
  return View<T>{(T*)t->data, t->ld, t->c, voxels(t), (int64_t)t->d * t->h * t->w};
}

static inline bool vec_ok(const b200_tensor* t, int vec) {
  return t->c % vec == 0 && t->ld % vec == 0 && ((uintptr_t)t->data % 16) == 0;
}

static inline unsigned grid_for(int64_t work, int threads, int waves = 8) {
  int64_t b = ceil_div(work, threads);
  int64_t cap = (int64_t)sm_count() * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* p, float (&f)[VEC]) {
  Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(p);
#pragma unroll
  for (int i = 0; i < VEC; ++i) f[i] = to_f<T>(v.v[i]);
}
template <typename T, int VEC>
__device__ __forceinline__ void store_vec(T* p, const float (&f)[VEC]) {
  Pack<T, VEC> v;
#pragma unroll
  for (int i = 0; i < VEC; ++i) v.v[i] = from_f<T>(f[i]);
  *reinterpret_cast<Pack<T, VEC>*>(p) = v;
}

#include "norm_fast.cuh"
#include "pack_batch.cuh"

// The job table travels as a kernel argument (by value, __grid_constant__): no device-side table to keep alive, nothing to upload,
// and a captured CUDA graph replays the launch as it is.  3.4 KB of the 4 KB argument space.
constexpr int kPackBatchMax = 48;
struct PackJobTable { PackJob j[kPackBatchMax]; };
template <typename T>
__global__ void __launch_bounds__(256) pack_batch_table_kernel(const __grid_constant__ PackJobTable tab, int n_jobs) {
  pack_batch_body<T>(tab.j, n_jobs);
}

// ------------------------------------------------------------------------------------------ channel sums
// grid = (chunks, N).  Block = rows x CV threads (CV = C/VEC channel vectors); every thread owns one channel
// vector and strides over the voxels of its chunk.  fp32 partials are flushed into fp64 every 32 voxels, the
// block result is reduced through shared memory and added to sums[n][c][0..1] with one fp64 atomic each.
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 4) channel_sums_kernel(View<const T> x, double* __restrict__ sums, int cv_count, int rows) {
  extern __shared__ double s_red[];  // rows * cv_count * VEC * 2
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int cv = tid % cv_count;
  const int row = tid / cv_count;
  if (row < rows) {
    // fp32 partials per thread (a few hundred voxels at most), fp64 across threads / blocks
    float s[VEC], s2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] = s2[i] = 0.f;
    const int64_t chunk = (x.spatial + gridDim.x - 1) / gridDim.x;
    const int64_t v0 = (int64_t)blockIdx.x * chunk;
    int64_t v1 = v0 + chunk;
    if (v1 > x.spatial) v1 = x.spatial;
    const T* base = x.p + (int64_t)n * x.spatial * x.ld + (int64_t)cv * VEC;
    // U independent 16-byte loads in flight per thread, kept packed until they are consumed (register budget: 4 blocks/SM)
    constexpr int U = 4;
    for (int64_t v = v0 + row; v < v1; v += (int64_t)U * rows) {
      Pack<T, VEC> pk[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v + (int64_t)u * rows < v1) pk[u] = *reinterpret_cast<const Pack<T, VEC>*>(base + (v + (int64_t)u * rows) * x.ld);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v + (int64_t)u * rows < v1) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float f = to_f<T>(pk[u].v[i]);
            s[i] += f;
            s2[i] = fmaf(f, f, s2[i]);
          }
        }
    }
    double* dst = s_red + ((int64_t)row * cv_count + cv) * VEC * 2;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      dst[2 * i] = (double)s[i];
      dst[2 * i + 1] = (double)s2[i];
    }
  }
  __syncthreads();
  const int items = cv_count * VEC * 2;
  for (int it = tid; it < items; it += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < rows; ++r) t += s_red[(int64_t)r * items + it];
    atomicAdd(&sums[((int64_t)n * x.c) * 2 + it], t);
  }
}

template <typename T>
static int launch_channel_sums(const b200_tensor* x, double* sums, cudaStream_t st) {
  constexpr int V = VecOf<T>::n;
  View<const T> xv{(const T*)x->data, x->ld, x->c, voxels(x), (int64_t)x->d * x->h * x->w};
  auto grid_of = [&](int rows) {
    int64_t chunks = ceil_div(xv.spatial, (int64_t)rows * 8);     // small tensors: many short blocks beat a few long ones
    int64_t cap = ceil_div((int64_t)sm_count() * 4, x->n);
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    return dim3((unsigned)chunks, x->n);
  };
  constexpr int VH = V;
  if (vec_ok(x, V) && x->c / VH <= 256) {
    int cvn = x->c / VH;
    int rows = 256 / cvn;
    int threads = ((rows * cvn + 31) / 32) * 32;
    size_t smem = sizeof(double) * rows * cvn * VH * 2;
    channel_sums_kernel<T, VH><<<grid_of(rows), threads, smem, st>>>(xv, sums, cvn, rows);
  } else {
    dim3 grid = grid_of(1);
    B200_CHECK_ARG(x->c <= 1024, "channel_sums: too many channels (%d)", x->c);
    int cvn = x->c;
    int rows = 256 / cvn;
    if (rows < 1) rows = 1;
    if (rows > 32) rows = 32;
    int threads = ((rows * cvn + 31) / 32) * 32;
    size_t smem = sizeof(double) * rows * cvn * 2;
    channel_sums_kernel<T, 1><<<grid, threads, smem, st>>>(xv, sums, cvn, rows);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ----------------------------------------------------------------------------------------- norm finalize
__global__ void norm_finalize_kernel(const double* __restrict__ sums, int n, int c, int groups, int64_t spatial,
                                     int batch_stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float eps, float* __restrict__ mean, float* __restrict__ rstd,
                                     float* __restrict__ scale, float* __restrict__ shift) {
  // one WARP per (n, g): the lanes stride over the group's channels (GroupNorm(8, C): up to 48 channels per group at the deep
  // levels -- one thread walking them through dependent double-precision loads was a 5-15 us kernel 17 times per pass)
  const int wid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (wid >= n * groups) return;
  const int ni = wid / groups, g = wid % groups;
  const int cpg = c / groups;
  double s = 0.0, s2 = 0.0;
  const int n0 = batch_stats ? 0 : ni, n1 = batch_stats ? n : ni + 1;
  for (int nn = n0; nn < n1; ++nn)
    for (int cc = g * cpg + lane; cc < (g + 1) * cpg; cc += 32) {
      s += sums[((int64_t)nn * c + cc) * 2];
      s2 += sums[((int64_t)nn * c + cc) * 2 + 1];
    }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const double m = (double)spatial * cpg * (n1 - n0);
  const double mu = s / m;
  double var = s2 / m - mu * mu;
  if (var < 0.0) var = 0.0;
  const float r = 1.0f / sqrtf((float)var + eps);
  if (lane == 0) {
    mean[wid] = (float)mu;
    rstd[wid] = r;
  }
  for (int cc = g * cpg + lane; cc < (g + 1) * cpg; cc += 32) {
    const float ga = gamma ? gamma[cc] : 1.f, be = beta ? beta[cc] : 0.f;
    const float sc = r * ga;
    scale[(int64_t)ni * c + cc] = sc;
    shift[(int64_t)ni * c + cc] = be - (float)mu * sc;
  }
}

// ----------------------------------------------------------------------------------- scale-shift-activation
template <typename TI, typename TO, int VEC>
__global__ void scale_shift_act_kernel(View<const TI> x, View<TO> y, const float* __restrict__ scale,
                                       const float* __restrict__ shift, int act) {
  const int cvn = x.c / VEC;
  const int64_t total = x.vox * cvn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t vox = i / cvn;
    int cv = (int)(i % cvn);
    int n = (int)(vox / x.spatial);
    float f[VEC];
    load_vec<TI, VEC>(x.p + vox * x.ld + cv * VEC, f);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float v = f[k];
      if (scale) v = fmaf(v, scale[(int64_t)n * x.c + cv * VEC + k], shift[(int64_t)n * x.c + cv * VEC + k]);
      f[k] = act_fwd(act, v);
    }
    store_vec<TO, VEC>(y.p + vox * y.ld + cv * VEC, f);
  }
}

// -------------------------------------------------------------------------------- norm+act backward, pass 1
// same thread layout as channel_sums; accumulates (sum g, sum g*(x - mean)) per (n,c)
template <typename T, int VEC, int ACT, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) norm_act_bwd_reduce_kernel(View<const T> x, View<const T> dy, const float* __restrict__ mean,
                                           const float* __restrict__ rstd, int groups,
                                           const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                           double* __restrict__ red, int cv_count, int rows) {
  extern __shared__ double s_red[];
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int cv = tid % cv_count;
  const int row = tid / cv_count;
  const int cpg = x.c / groups;
  if (row < rows) {
    // ypre = x*ka + kb with ka = rs*ga, kb = be - mu*rs*ga.  Accumulates (sum g, sum g*(x - mu)): the CENTRED second moment,
    // so that sum g*xhat = rs * it without the cancellation of rs*sum(g*x) - mu*rs*sum(g) (fp32 partials over thousands of
    // voxels per thread: with |mu| of a few sigma that difference lost 3-4 digits -- input gradients of the 64^3 Attention
    // U-Net were 3e-2 off the CPU reference, 1e-3 after centring).
    float ka[VEC], kb[VEC], mu_c[VEC], s[VEC], s2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      int c = cv * VEC + i;
      int g = c / cpg;
      float mu = mean[(int64_t)n * groups + g], r = rstd[(int64_t)n * groups + g];
      float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      ka[i] = r * ga;
      kb[i] = be - mu * r * ga;
      mu_c[i] = mu;
      s[i] = s2[i] = 0.f;
    }
    const int64_t chunk = (x.spatial + gridDim.x - 1) / gridDim.x;
    const int64_t v0 = (int64_t)blockIdx.x * chunk;
    int64_t v1 = v0 + chunk;
    if (v1 > x.spatial) v1 = x.spatial;
    const T* xb = x.p + (int64_t)n * x.spatial * x.ld + (int64_t)cv * VEC;
    const T* db = dy.p + (int64_t)n * x.spatial * dy.ld + (int64_t)cv * VEC;
    // U packed 16-byte loads in flight per thread and tensor
    for (int64_t v = v0 + row; v < v1; v += (int64_t)U * rows) {
      Pack<T, VEC> px[U], pd[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v + (int64_t)u * rows < v1) {
          px[u] = *reinterpret_cast<const Pack<T, VEC>*>(xb + (v + (int64_t)u * rows) * x.ld);
          pd[u] = *reinterpret_cast<const Pack<T, VEC>*>(db + (v + (int64_t)u * rows) * dy.ld);
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v + (int64_t)u * rows < v1) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float fx = to_f<T>(px[u].v[i]);
            const float g = to_f<T>(pd[u].v[i]) * act_grad_t<ACT>(act, fmaf(fx, ka[i], kb[i]));
            s[i] += g;
            s2[i] = fmaf(g, fx - mu_c[i], s2[i]);
          }
        }
    }
    double* dst = s_red + ((int64_t)row * cv_count + cv) * VEC * 2;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      dst[2 * i] = (double)s[i];
      dst[2 * i + 1] = (double)s2[i];
    }
  }
  __syncthreads();
  const int items = cv_count * VEC * 2;
  for (int it = tid; it < items; it += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < rows; ++r) t += s_red[(int64_t)r * items + it];
    atomicAdd(&red[((int64_t)n * x.c) * 2 + it], t);
  }
}

// tiny: per (n, g) coefficients + dgamma/dbeta
__global__ void norm_bwd_finalize_kernel(const double* __restrict__ red, const float* __restrict__ mean,
                                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, int n, int c, int groups, int64_t spatial,
                                         int batch_stats, float* __restrict__ coef, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta, const double* __restrict__ xsums = nullptr,
                                         float* __restrict__ dxsum = nullptr) {
  // red[n][c] = (S1 = sum g, Sgc = sum g*(x - mean));  S2 = sum g*xhat = rstd*Sgc
  // one WARP per (n, g) for the coefficients (lanes over the group's channels), one THREAD per channel for dgamma / dbeta
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int wid = idx >> 5, lane = threadIdx.x & 31;
  const int cpg = c / groups;
  if (wid < n * groups) {
    const int ni = wid / groups, g = wid % groups;
    const int n0 = batch_stats ? 0 : ni, n1 = batch_stats ? n : ni + 1;
    double a = 0.0, b = 0.0;
    for (int nn = n0; nn < n1; ++nn) {
      const double r = (double)rstd[(int64_t)nn * groups + g];
      for (int cc = g * cpg + lane; cc < (g + 1) * cpg; cc += 32) {
        const double ga = gamma ? (double)gamma[cc] : 1.0;
        const double s1 = red[((int64_t)nn * c + cc) * 2], sgc = red[((int64_t)nn * c + cc) * 2 + 1];
        a += ga * s1;
        b += ga * (r * sgc);
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const double m = (double)spatial * cpg * (n1 - n0);
    // with ypre = x*k0 + B and g = dy*act'(ypre):  dx = g*k0 - x*P - Q
    const float r = rstd[wid], mu = mean[wid];
    const float k1 = (float)(r * (a / m)), k2 = (float)(r * (b / m));
    for (int cc = g * cpg + lane; cc < (g + 1) * cpg; cc += 32) {
      const float ga = gamma ? gamma[cc] : 1.f, be = beta ? beta[cc] : 0.f;
      float* o = coef + ((int64_t)ni * c + cc) * 4;
      const float k0 = r * ga, kp = r * k2, kq = k1 - mu * r * k2;
      o[0] = k0;
      o[1] = be - mu * r * ga;
      o[2] = kp;
      o[3] = kq;
      // dxsum[c] += sum over this sample's voxels of dx = g*k0 - x*P - Q, from the sums alone (xsums[n][c][0] = sum x of the
      // forward statistics): the bias gradient of a convolution whose output feeds this normalisation only -- no pass over dx
      if (dxsum && xsums)
        atomicAdd(&dxsum[cc], (float)((double)k0 * red[((int64_t)ni * c + cc) * 2] - (double)kp * xsums[((int64_t)ni * c + cc) * 2] -
                                      (double)kq * (double)spatial));
    }
  }
  if (idx < c) {
    int g = idx / cpg;
    double sg = 0.0, sb = 0.0;
    for (int nn = 0; nn < n; ++nn) {
      const double r = (double)rstd[(int64_t)nn * groups + g];
      double s1 = red[((int64_t)nn * c + idx) * 2], sgc = red[((int64_t)nn * c + idx) * 2 + 1];
      sb += s1;
      sg += r * sgc;
    }
    if (dgamma) dgamma[idx] += (float)sg;
    if (dbeta) dbeta[idx] += (float)sb;
  }
}

// xsum[ci] += sum_co w[co][ci] * dysum[co]: the channel sums of a gradient pass through a pointwise convolution's input
// gradient as they are (dx = W^T dy voxel by voxel), so the sums of `x.grad` stay known after a 1x1 shortcut accumulates into it
__global__ void sums_through_pointwise_kernel(const float* __restrict__ w, const float* __restrict__ dysum, float* __restrict__ xsum,
                                              int cout, int cin) {
  const int ci = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;   // one warp per input channel
  if (ci >= cin) return;
  double acc = 0.0;
  for (int co = lane; co < cout; co += 32) acc += (double)w[(int64_t)co * cin + ci] * (double)dysum[co];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) xsum[ci] += (float)acc;
}

// element-wise fallback (any channel count): coefficients re-read per element
template <typename T>
__global__ void norm_act_bwd_apply_kernel(View<const T> x, View<const T> dy, View<T> dx, int act, const float* __restrict__ coef,
                                          int accumulate) {
  const int64_t total = x.vox * x.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t vox = i / x.c;
    int c = (int)(i % x.c);
    int n = (int)(vox / x.spatial);
    const float* cf = coef + ((int64_t)n * x.c + c) * 4;
    float xv = to_f<T>(x.p[vox * x.ld + c]);
    float g = to_f<T>(dy.p[vox * dy.ld + c]) * act_grad(act, fmaf(xv, cf[0], cf[1]));
    float r = g * cf[0] - xv * cf[2] - cf[3];
    T* o = dx.p + vox * dx.ld + c;
    *o = from_f<T>(accumulate ? to_f<T>(*o) + r : r);
  }
}

// vectorised: grid = (chunks, N), block = rows x cvn threads; every thread keeps the coefficients of its 8 (4) channels
// in registers, so the loop body is 2-3 16-byte loads, the activation derivative and one 16-byte store
template <typename T, int VEC, int ACT, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) norm_act_bwd_apply_rows_kernel(View<const T> x, View<const T> dy, View<T> dx, int act,
                                               const float* __restrict__ coef, int accumulate, int cvn, int rows) {
  const int n = blockIdx.y;
  const int cv = threadIdx.x % cvn, row = threadIdx.x / cvn;
  if (row >= rows) return;
  float k0[VEC], kb[VEC], kp[VEC], kq[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float* cf = coef + ((int64_t)n * x.c + cv * VEC + i) * 4;
    k0[i] = cf[0]; kb[i] = cf[1]; kp[i] = cf[2]; kq[i] = cf[3];
  }
  const T* xb = x.p + (int64_t)n * x.spatial * x.ld + cv * VEC;
  const T* db = dy.p + (int64_t)n * x.spatial * dy.ld + cv * VEC;
  T* ob = dx.p + (int64_t)n * x.spatial * dx.ld + cv * VEC;
  // U independent 16-byte loads in flight per thread and tensor
  const int64_t stride = (int64_t)gridDim.x * rows;
  for (int64_t v0 = (int64_t)blockIdx.x * rows + row; v0 < x.spatial; v0 += U * stride) {
    Pack<T, VEC> px[U], pd[U], po[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < x.spatial) {
        px[u] = *reinterpret_cast<const Pack<T, VEC>*>(xb + v * x.ld);
        pd[u] = *reinterpret_cast<const Pack<T, VEC>*>(db + v * dy.ld);
        if (accumulate) po[u] = *reinterpret_cast<const Pack<T, VEC>*>(ob + v * dx.ld);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < x.spatial) {
        Pack<T, VEC> out;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float fx = to_f<T>(px[u].v[i]);
          const float g = to_f<T>(pd[u].v[i]) * act_grad_t<ACT>(act, fmaf(fx, k0[i], kb[i]));
          float r = g * k0[i] - fx * kp[i] - kq[i];
          if (accumulate) r += to_f<T>(po[u].v[i]);
          out.v[i] = from_f<T>(r);
        }
        *reinterpret_cast<Pack<T, VEC>*>(ob + v * dx.ld) = out;
      }
    }
  }
}

template <typename T, int VEC, int ACT>
__global__ void scale_shift_act_rows_kernel(View<const T> x, View<T> y, const float* __restrict__ scale,
                                            const float* __restrict__ shift, int act, int cvn, int rows) {
  const int n = blockIdx.y;
  const int cv = threadIdx.x % cvn, row = threadIdx.x / cvn;
  if (row >= rows) return;
  float sc[VEC], sh[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    sc[i] = scale[(int64_t)n * x.c + cv * VEC + i];
    sh[i] = shift[(int64_t)n * x.c + cv * VEC + i];
  }
  const T* xb = x.p + (int64_t)n * x.spatial * x.ld + cv * VEC;
  T* yb = y.p + (int64_t)n * x.spatial * y.ld + cv * VEC;
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * rows;
  for (int64_t v0 = (int64_t)blockIdx.x * rows + row; v0 < x.spatial; v0 += U * stride) {
    Pack<T, VEC> px[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < x.spatial) px[u] = *reinterpret_cast<const Pack<T, VEC>*>(xb + v * x.ld);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < x.spatial) {
        Pack<T, VEC> out;
#pragma unroll
        for (int i = 0; i < VEC; ++i) out.v[i] = from_f<T>(act_fwd_t<ACT>(act, fmaf(to_f<T>(px[u].v[i]), sc[i], sh[i])));
        *reinterpret_cast<Pack<T, VEC>*>(yb + v * y.ld) = out;
      }
    }
  }
}

template <typename T, int VEC>
__global__ void act_bwd_kernel(View<const T> x, View<const T> dy, View<T> dx, int act, int accumulate) {
  const int cvn = x.c / VEC;
  const int64_t total = x.vox * cvn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t vox = i / cvn;
    int cv = (int)(i % cvn);
    float fx[VEC], fd[VEC], fo[VEC];
    load_vec<T, VEC>(x.p + vox * x.ld + cv * VEC, fx);
    load_vec<T, VEC>(dy.p + vox * dy.ld + cv * VEC, fd);
    if (accumulate) load_vec<T, VEC>(dx.p + vox * dx.ld + cv * VEC, fo);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float r = fd[k] * act_grad(act, fx[k]);
      fo[k] = accumulate ? fo[k] + r : r;
    }
    store_vec<T, VEC>(dx.p + vox * dx.ld + cv * VEC, fo);
  }
}

// ------------------------------------------------------------------------------------------------ max-pool
struct PoolGeom {
  int n, d, h, w, c;     // input
  int od, oh, ow;        // output
  int pd, ph, pw;
};

template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, int64_t ldx, T* __restrict__ y, int64_t ldy, PoolGeom g) {
  const int64_t total = (int64_t)g.n * g.od * g.oh * g.ow * g.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    int c = (int)(t % g.c); t /= g.c;
    int ox = (int)(t % g.ow); t /= g.ow;
    int oy = (int)(t % g.oh); t /= g.oh;
    int oz = (int)(t % g.od); t /= g.od;
    int n = (int)t;
    float m = -INFINITY;
    for (int a = 0; a < g.pd; ++a)
      for (int b = 0; b < g.ph; ++b)
        for (int e = 0; e < g.pw; ++e) {
          int64_t vox = (((int64_t)n * g.d + oz * g.pd + a) * g.h + oy * g.ph + b) * g.w + ox * g.pw + e;
          float v = to_f<T>(x[vox * ldx + c]);
          if (v > m || v != v) m = v;
        }
    int64_t ov = (((int64_t)n * g.od + oz) * g.oh + oy) * g.ow + ox;
    y[ov * ldy + c] = from_f<T>(m);
  }
}

// one thread per INPUT element: finds the first maximum of its window in (d,h,w) scan order
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                                   T* __restrict__ dx, int64_t lddx, PoolGeom g, int accumulate) {
  const int64_t total = (int64_t)g.n * g.d * g.h * g.w * g.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    int c = (int)(t % g.c); t /= g.c;
    int xx = (int)(t % g.w); t /= g.w;
    int yy = (int)(t % g.h); t /= g.h;
    int zz = (int)(t % g.d); t /= g.d;
    int n = (int)t;
    int oz = zz / g.pd, oy = yy / g.ph, ox = xx / g.pw;
    float r = 0.f;
    if (oz < g.od && oy < g.oh && ox < g.ow) {
      float m = -INFINITY;
      int arg = 0, k = 0;
      for (int a = 0; a < g.pd; ++a)
        for (int b = 0; b < g.ph; ++b)
          for (int e = 0; e < g.pw; ++e, ++k) {
            int64_t vox = (((int64_t)n * g.d + oz * g.pd + a) * g.h + oy * g.ph + b) * g.w + ox * g.pw + e;
            float v = to_f<T>(x[vox * ldx + c]);
            if (v > m || v != v) { m = v; arg = k; }
          }
      int mine = ((zz - oz * g.pd) * g.ph + (yy - oy * g.ph)) * g.pw + (xx - ox * g.pw);
      if (mine == arg) {
        int64_t ov = (((int64_t)n * g.od + oz) * g.oh + oy) * g.ow + ox;
        r = to_f<T>(dy[ov * lddy + c]);
      }
    }
    int64_t iv = (((int64_t)n * g.d + zz) * g.h + yy) * g.w + xx;
    T* o = dx + iv * lddx + c;
    *o = from_f<T>(accumulate ? to_f<T>(*o) + r : r);
  }
}

// Vectorised pooling: one thread per pooled voxel and channel vector; every input element is read once and every
// gradient element written once (the scalar kernels above stay as the fallback for odd shapes / channel counts).
template <typename T, int VEC>
__global__ void maxpool_fwd_vec_kernel(const T* __restrict__ x, int64_t ldx, T* __restrict__ y, int64_t ldy, PoolGeom g) {
  const int cvn = g.c / VEC;
  const int64_t total = (int64_t)g.n * g.od * g.oh * g.ow * cvn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    int cv = (int)(t % cvn); t /= cvn;
    int ox = (int)(t % g.ow); t /= g.ow;
    int oy = (int)(t % g.oh); t /= g.oh;
    int oz = (int)(t % g.od); t /= g.od;
    int n = (int)t;
    float m[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) m[k] = -INFINITY;
    for (int a = 0; a < g.pd; ++a)
      for (int b = 0; b < g.ph; ++b)
        for (int e = 0; e < g.pw; ++e) {
          int64_t vox = (((int64_t)n * g.d + oz * g.pd + a) * g.h + oy * g.ph + b) * g.w + ox * g.pw + e;
          float f[VEC];
          load_vec<T, VEC>(x + vox * ldx + cv * VEC, f);
#pragma unroll
          for (int k = 0; k < VEC; ++k)
            if (f[k] > m[k] || f[k] != f[k]) m[k] = f[k];
        }
    int64_t ov = (((int64_t)n * g.od + oz) * g.oh + oy) * g.ow + ox;
    store_vec<T, VEC>(y + ov * ldy + cv * VEC, m);
  }
}

template <typename T, int VEC, int MAXW>
__global__ void maxpool_bwd_vec_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                                       T* __restrict__ dx, int64_t lddx, PoolGeom g, int accumulate) {
  const int cvn = g.c / VEC;
  const int64_t total = (int64_t)g.n * g.od * g.oh * g.ow * cvn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    int cv = (int)(t % cvn); t /= cvn;
    int ox = (int)(t % g.ow); t /= g.ow;
    int oy = (int)(t % g.oh); t /= g.oh;
    int oz = (int)(t % g.od); t /= g.od;
    int n = (int)t;
    float m[VEC];
    int arg[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { m[k] = -INFINITY; arg[k] = 0; }
    int w = 0;
    for (int a = 0; a < g.pd; ++a)
      for (int b = 0; b < g.ph; ++b)
        for (int e = 0; e < g.pw; ++e, ++w) {
          int64_t vox = (((int64_t)n * g.d + oz * g.pd + a) * g.h + oy * g.ph + b) * g.w + ox * g.pw + e;
          float f[VEC];
          load_vec<T, VEC>(x + vox * ldx + cv * VEC, f);
#pragma unroll
          for (int k = 0; k < VEC; ++k)
            if (f[k] > m[k] || f[k] != f[k]) { m[k] = f[k]; arg[k] = w; }
        }
    int64_t ov = (((int64_t)n * g.od + oz) * g.oh + oy) * g.ow + ox;
    float gy[VEC];
    load_vec<T, VEC>(dy + ov * lddy + cv * VEC, gy);
    w = 0;
    for (int a = 0; a < g.pd; ++a)
      for (int b = 0; b < g.ph; ++b)
        for (int e = 0; e < g.pw; ++e, ++w) {
          int64_t vox = (((int64_t)n * g.d + oz * g.pd + a) * g.h + oy * g.ph + b) * g.w + ox * g.pw + e;
          T* o = dx + vox * lddx + cv * VEC;
          float r[VEC];
          if (accumulate) load_vec<T, VEC>(o, r);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            float v = (arg[k] == w) ? gy[k] : 0.f;
            r[k] = accumulate ? r[k] + v : v;
          }
          store_vec<T, VEC>(o, r);
        }
  }
}

// Compile-time window (the U-Net family pools 2x2x2, 1x2x2 or 2x2): every load of a thread -- the PD*PH*PW window vectors,
// the pooled gradient and, when accumulating, the old gradient vectors -- is issued before the first compare, so a thread
// keeps 8-17 independent 16-byte requests in flight instead of one (the runtime-window loops above cannot be unrolled:
// 0.44 + 0.19 ms per cfg-2 step against a 0.15 + 0.05 ms HBM floor).  Same scan order and NaN rule as the kernels above.
// IDX = uint32_t when the thread count fits 31 bits: the five div / mod pairs of the index decode cost ~100 instructions
// each in 64-bit arithmetic, more than the memory instructions of the thread.
template <typename T, int VEC, int PD, int PH, int PW, typename IDX>
__global__ void __launch_bounds__(256) maxpool_fwd_win_kernel(const T* __restrict__ x, int64_t ldx, T* __restrict__ y, int64_t ldy,
                                                              PoolGeom g) {
  constexpr int NW = PD * PH * PW;
  const IDX cvn = (IDX)(g.c / VEC);
  const IDX total = (IDX)((int64_t)g.n * g.od * g.oh * g.ow * (g.c / VEC));
  const IDX stride = (IDX)gridDim.x * (IDX)blockDim.x;
  for (IDX i = (IDX)blockIdx.x * (IDX)blockDim.x + threadIdx.x; i < total; i += stride) {
    IDX t = i;
    const int cv = (int)(t % cvn); t /= cvn;
    const int ox = (int)(t % (IDX)g.ow); t /= (IDX)g.ow;
    const int oy = (int)(t % (IDX)g.oh); t /= (IDX)g.oh;
    const int oz = (int)(t % (IDX)g.od); t /= (IDX)g.od;
    const int n = (int)t;
    const T* base = x + ((((int64_t)n * g.d + oz * PD) * g.h + oy * PH) * g.w + ox * PW) * ldx + cv * VEC;
    Pack<T, VEC> p[NW];
#pragma unroll
    for (int a = 0; a < PD; ++a)
#pragma unroll
      for (int b = 0; b < PH; ++b)
#pragma unroll
        for (int e = 0; e < PW; ++e)
          p[(a * PH + b) * PW + e] = *reinterpret_cast<const Pack<T, VEC>*>(base + (((int64_t)a * g.h + b) * g.w + e) * ldx);
    float m[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float f = to_f<T>(p[w].v[k]);
        if (f > m[k] || f != f) m[k] = f;
      }
    const int64_t ov = (((int64_t)n * g.od + oz) * g.oh + oy) * g.ow + ox;
    store_vec<T, VEC>(y + ov * ldy + cv * VEC, m);
  }
}

template <typename T, int VEC, int PD, int PH, int PW, bool ACC, typename IDX>
__global__ void __launch_bounds__(256) maxpool_bwd_win_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy,
                                                              int64_t lddy, T* dx, int64_t lddx, const T* dsrc, int64_t ldsrc, PoolGeom g) {
  constexpr int NW = PD * PH * PW;
  const IDX cvn = (IDX)(g.c / VEC);
  const IDX total = (IDX)((int64_t)g.n * g.od * g.oh * g.ow * (g.c / VEC));
  const IDX stride = (IDX)gridDim.x * (IDX)blockDim.x;
  for (IDX i = (IDX)blockIdx.x * (IDX)blockDim.x + threadIdx.x; i < total; i += stride) {
    IDX t = i;
    const int cv = (int)(t % cvn); t /= cvn;
    const int ox = (int)(t % (IDX)g.ow); t /= (IDX)g.ow;
    const int oy = (int)(t % (IDX)g.oh); t /= (IDX)g.oh;
    const int oz = (int)(t % (IDX)g.od); t /= (IDX)g.od;
    const int n = (int)t;
    const int64_t vox0 = (((int64_t)n * g.d + oz * PD) * g.h + oy * PH) * g.w + ox * PW;
    const T* xb = x + vox0 * ldx + cv * VEC;
    T* db = dx + vox0 * lddx + cv * VEC;
    // ACC: the routed gradient is added to `dsrc` (== dx for the in-place form; another tensor, possibly with another voxel pitch,
    // when the sum is handed on as a dense tensor while the running gradient lives in a channel slice of a concat buffer)
    const T* sb = ACC ? dsrc + vox0 * ldsrc + cv * VEC : nullptr;
    const int64_t ov = (((int64_t)n * g.od + oz) * g.oh + oy) * g.ow + ox;
    Pack<T, VEC> p[NW], old[NW];
#pragma unroll
    for (int a = 0; a < PD; ++a)
#pragma unroll
      for (int b = 0; b < PH; ++b)
#pragma unroll
        for (int e = 0; e < PW; ++e) {
          const int64_t off = ((int64_t)a * g.h + b) * g.w + e;
          p[(a * PH + b) * PW + e] = *reinterpret_cast<const Pack<T, VEC>*>(xb + off * ldx);
          if (ACC) old[(a * PH + b) * PW + e] = *reinterpret_cast<const Pack<T, VEC>*>(sb + off * ldsrc);
        }
    const Pack<T, VEC> pg = *reinterpret_cast<const Pack<T, VEC>*>(dy + ov * lddy + cv * VEC);
    float m[VEC];
    int arg[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { m[k] = -INFINITY; arg[k] = 0; }
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float f = to_f<T>(p[w].v[k]);
        if (f > m[k] || f != f) { m[k] = f; arg[k] = w; }
      }
#pragma unroll
    for (int a = 0; a < PD; ++a)
#pragma unroll
      for (int b = 0; b < PH; ++b)
#pragma unroll
        for (int e = 0; e < PW; ++e) {
          const int w = (a * PH + b) * PW + e;
          Pack<T, VEC> out;
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float v = (arg[k] == w) ? to_f<T>(pg.v[k]) : 0.f;
            out.v[k] = from_f<T>(ACC ? to_f<T>(old[w].v[k]) + v : v);
          }
          *reinterpret_cast<Pack<T, VEC>*>(db + (((int64_t)a * g.h + b) * g.w + e) * lddx) = out;
        }
  }
}

// B200_POOL_WIN=0 falls back to the runtime-window kernels
static inline bool pool_win_enabled() {
  static const bool on = !(getenv("B200_POOL_WIN") && atoi(getenv("B200_POOL_WIN")) == 0);
  return on;
}

template <typename T, int V, int PD, typename IDX>
static void launch_pool_bwd_win2(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx, const b200_tensor* dsrc,
                                 const PoolGeom& g, int accumulate, cudaStream_t st) {
  const unsigned grid = grid_for(voxels(dy) * (x->c / V), 256);
  const T* xp = (const T*)x->data;
  const T* gp = (const T*)dy->data;
  T* dp = (T*)dx->data;
  const T* sp = dsrc ? (const T*)dsrc->data : dp;
  const int64_t lds = dsrc ? dsrc->ld : dx->ld;
  if (accumulate) maxpool_bwd_win_kernel<T, V, PD, 2, 2, true, IDX><<<grid, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dp, dx->ld, sp, lds, g);
  else maxpool_bwd_win_kernel<T, V, PD, 2, 2, false, IDX><<<grid, 256, 0, st>>>(xp, x->ld, gp, dy->ld, dp, dx->ld, sp, lds, g);
}

// 32-bit index decode whenever the (thread count + one grid stride) stays below 2^31
static inline bool pool_idx32(int64_t threads) { return threads + (int64_t)sm_count() * 8 * 256 < ((int64_t)1 << 31); }

template <typename T, int V>
static void launch_pool_bwd_win(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx, const b200_tensor* dsrc,
                                const PoolGeom& g, int pd, int accumulate, cudaStream_t st) {
  const bool small = pool_idx32(voxels(dy) * (x->c / V));
  if (pd == 2) {
    if (small) launch_pool_bwd_win2<T, V, 2, uint32_t>(x, dy, dx, dsrc, g, accumulate, st);
    else launch_pool_bwd_win2<T, V, 2, int64_t>(x, dy, dx, dsrc, g, accumulate, st);
  } else {
    if (small) launch_pool_bwd_win2<T, V, 1, uint32_t>(x, dy, dx, dsrc, g, accumulate, st);
    else launch_pool_bwd_win2<T, V, 1, int64_t>(x, dy, dx, dsrc, g, accumulate, st);
  }
}

template <typename T, int V>
static void launch_pool_fwd_win(const b200_tensor* x, const b200_tensor* y, const PoolGeom& g, int pd, cudaStream_t st) {
  const int64_t threads = voxels(y) * (y->c / V);
  const unsigned grid = grid_for(threads, 256);
  const T* xp = (const T*)x->data;
  T* yp = (T*)y->data;
  if (pool_idx32(threads)) {
    if (pd == 2) maxpool_fwd_win_kernel<T, V, 2, 2, 2, uint32_t><<<grid, 256, 0, st>>>(xp, x->ld, yp, y->ld, g);
    else maxpool_fwd_win_kernel<T, V, 1, 2, 2, uint32_t><<<grid, 256, 0, st>>>(xp, x->ld, yp, y->ld, g);
  } else {
    if (pd == 2) maxpool_fwd_win_kernel<T, V, 2, 2, 2, int64_t><<<grid, 256, 0, st>>>(xp, x->ld, yp, y->ld, g);
    else maxpool_fwd_win_kernel<T, V, 1, 2, 2, int64_t><<<grid, 256, 0, st>>>(xp, x->ld, yp, y->ld, g);
  }
}

// ----------------------------------------------------------------------------------------------- binary ops
template <typename TA, typename TB, typename TY>
__global__ void binary_kernel(View<const TA> a, View<const TB> b, View<TY> y, int op, int b_bcast) {
  const int64_t total = a.vox * a.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t vox = i / a.c;
    int c = (int)(i % a.c);
    float va = to_f<TA>(a.p[vox * a.ld + c]);
    float vb = 0.f;
    if (op != 3 && op != 4) vb = to_f<TB>(b.p[vox * b.ld + (b_bcast ? 0 : c)]);
    float r;
    switch (op) {
      case 0: r = va + vb; break;
      case 1: r = va * vb; break;
      case 2: r = fmaxf(va + vb, 0.f); break;
      case 3: r = va; break;
      default: r = 1.f / (1.f + expf(-va)); break;
    }
    y.p[vox * y.ld + c] = from_f<TY>(r);
  }
}

// 16-byte form (no channel broadcast): thread = (voxel, 8-channel vector), two vectors in flight
template <typename T, int VEC, int OP>
__global__ void __launch_bounds__(256, 4) binary_vec_kernel(View<const T> a, View<const T> b, View<T> y) {
  const int cvn = a.c / VEC;
  const int64_t total = a.vox * cvn;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 2;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
    Pack<T, VEC> pa[U], pb[U];
    int64_t vox[U];
    int cv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      vox[u] = i / cvn;
      cv[u] = (int)(i - vox[u] * cvn) * VEC;
      if (i < total) {
        pa[u] = *reinterpret_cast<const Pack<T, VEC>*>(a.p + vox[u] * a.ld + cv[u]);
        if (OP != 3 && OP != 4) pb[u] = *reinterpret_cast<const Pack<T, VEC>*>(b.p + vox[u] * b.ld + cv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * stride >= total) continue;
      Pack<T, VEC> out;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float va = to_f<T>(pa[u].v[k]);
        const float vb = (OP != 3 && OP != 4) ? to_f<T>(pb[u].v[k]) : 0.f;
        float r;
        if (OP == 0) r = va + vb;
        else if (OP == 1) r = va * vb;
        else if (OP == 2) r = fmaxf(va + vb, 0.f);
        else if (OP == 3) r = va;
        else r = 1.f / (1.f + expf(-va));
        out.v[k] = from_f<T>(r);
      }
      *reinterpret_cast<Pack<T, VEC>*>(y.p + vox[u] * y.ld + cv[u]) = out;
    }
  }
}

// out = psi * x (psi has 1 channel): dpsi[vox] = sum_c dout*x ; dx (+)= dout*psi
template <typename T>
__global__ void gate_bwd_kernel(View<const T> x, View<const T> psi, View<const T> dout, View<T> dpsi, View<T> dx,
                                int accumulate) {
  for (int64_t vox = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vox < x.vox; vox += (int64_t)gridDim.x * blockDim.x) {
    float p = to_f<T>(psi.p[vox * psi.ld]);
    float s = 0.f;
    for (int c = 0; c < x.c; ++c) {
      float d = to_f<T>(dout.p[vox * dout.ld + c]);
      s = fmaf(d, to_f<T>(x.p[vox * x.ld + c]), s);
      T* o = dx.p + vox * dx.ld + c;
      float r = d * p;
      *o = from_f<T>(accumulate ? to_f<T>(*o) + r : r);
    }
    dpsi.p[vox * dpsi.ld] = from_f<T>(s);
  }
}

template <typename T>
__global__ void relu_mask_bwd_kernel(View<const T> y, View<const T> dy, View<T> da) {
  const int64_t total = y.vox * y.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t vox = i / y.c;
    int c = (int)(i % y.c);
    float v = to_f<T>(y.p[vox * y.ld + c]);
    float d = to_f<T>(dy.p[vox * dy.ld + c]);
    da.p[vox * da.ld + c] = from_f<T>(v > 0.f ? d : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------- losses
__device__ __forceinline__ void block_atomic_add(double v, double* dst) {
  __shared__ double s_w[32];
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s_w[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < (int)((blockDim.x + 31) / 32)) ? s_w[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) atomicAdd(dst, t);
  }
  __syncthreads();
}

template <typename T>
__global__ void bce_logits_kernel(View<const T> z, const float* __restrict__ target, double* __restrict__ loss_sum,
                                  View<T> dz, float grad_scale) {
  const int64_t total = z.vox * z.c;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t vox = i / z.c;
    int c = (int)(i % z.c);
    float v = to_f<T>(z.p[vox * z.ld + c]);
    float t = target[i];
    // max(z,0) - z*t + log1p(exp(-|z|))
    float l = fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v)));
    acc += (double)l;
    if (dz.p) {
      float s = 1.f / (1.f + expf(-v));
      dz.p[vox * dz.ld + c] = from_f<T>((s - t) * grad_scale);
    }
  }
  block_atomic_add(acc, loss_sum);
}

// the same on dense 16-bit tensors (ld == c, element count a multiple of 8): eight logits per 16-byte load, one exponential per
// element (loss and sigmoid share exp(-|z|)), float partial sums per thread (a few dozen terms) folded in double
template <typename T>
__global__ void __launch_bounds__(256) bce_logits_dense_kernel(const T* __restrict__ z, const float* __restrict__ target,
                                                               double* __restrict__ loss_sum, T* __restrict__ dz, float grad_scale,
                                                               int64_t total8) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const Pack<T, 8> pz = *reinterpret_cast<const Pack<T, 8>*>(z + i * 8);
    const Pack<float, 4> t0 = *reinterpret_cast<const Pack<float, 4>*>(target + i * 8);
    const Pack<float, 4> t1 = *reinterpret_cast<const Pack<float, 4>*>(target + i * 8 + 4);
    const float t[8] = {t0.v[0], t0.v[1], t0.v[2], t0.v[3], t1.v[0], t1.v[1], t1.v[2], t1.v[3]};
    Pack<T, 8> out;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = to_f<T>(pz.v[k]);
      const float e = expf(-fabsf(v));
      acc += fmaxf(v, 0.f) - v * t[k] + log1pf(e);            // max(z,0) - z*t + log1p(exp(-|z|))
      const float s = (v >= 0.f ? 1.f : e) / (1.f + e);       // sigmoid(z)
      out.v[k] = from_f<T>((s - t[k]) * grad_scale);
    }
    if (dz) *reinterpret_cast<Pack<T, 8>*>(dz + i * 8) = out;
  }
  block_atomic_add((double)acc, loss_sum);
}

// target dense (vox, 2C): first C = target, last C = mask
template <typename T>
__global__ void n2v_mse_kernel(View<const T> y, const float* __restrict__ target, double* __restrict__ sums,
                               View<T> dy, float grad_scale, int mode) {
  const int64_t total = y.vox * y.c;
  double a = 0.0, b = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t vox = i / y.c;
    int c = (int)(i % y.c);
    float v = to_f<T>(y.p[vox * y.ld + c]);
    float t = target[vox * 2 * y.c + c];
    float m = target[vox * 2 * y.c + y.c + c];
    float e = t - v * m;
    if (mode != 1) {
      a += (double)e * e;
      b += (double)m;
    }
    if (mode != 0) dy.p[vox * dy.ld + c] = from_f<T>(-2.f * e * m * grad_scale);
  }
  if (mode != 1) {
    block_atomic_add(a, sums);
    block_atomic_add(b, sums + 1);
  }
}

// softmax CE over channels; one thread per voxel.  torch.nn.CrossEntropyLoss(ignore_index) semantics (reference metrics.py:546):
// voxels labelled `ignore_index` contribute neither loss nor gradient and the mean runs over the others -- sums[0] += loss,
// sums[1] += number of counted voxels.  A label outside [0, C) that is not the ignore value would be a device-side assert in
// torch; here the voxel is skipped the same way and counted in sums[2] so the host can raise.
template <typename T>
__global__ void softmax_ce_kernel(View<const T> z, const int64_t* __restrict__ target, double* __restrict__ sums,
                                  View<T> dz, float grad_scale, int64_t ignore_index) {
  double acc = 0.0, cnt = 0.0, bad = 0.0;
  for (int64_t vox = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vox < z.vox; vox += (int64_t)gridDim.x * blockDim.x) {
    const T* p = z.p + vox * z.ld;
    const int64_t t = target[vox];
    const bool valid = t >= 0 && t < z.c;
    if (!valid) {
      if (t != ignore_index) bad += 1.0;
      if (dz.p) {
        T* o = dz.p + vox * dz.ld;
        for (int c = 0; c < z.c; ++c) o[c] = from_f<T>(0.f);
      }
      continue;
    }
    float mx = -INFINITY;
    for (int c = 0; c < z.c; ++c) mx = fmaxf(mx, to_f<T>(p[c]));
    float se = 0.f;
    for (int c = 0; c < z.c; ++c) se += expf(to_f<T>(p[c]) - mx);
    float lse = mx + logf(se);
    acc += (double)(lse - to_f<T>(p[t]));
    cnt += 1.0;
    if (dz.p) {
      T* o = dz.p + vox * dz.ld;
      for (int c = 0; c < z.c; ++c) {
        float sm = expf(to_f<T>(p[c]) - lse);
        o[c] = from_f<T>((sm - (c == (int)t ? 1.f : 0.f)) * grad_scale);
      }
    }
  }
  block_atomic_add(acc, sums);
  block_atomic_add(cnt, sums + 1);
  block_atomic_add(bad, sums + 2);
}

template <typename TI, typename TO>
__global__ void softmax_channels_kernel(View<const TI> x, View<TO> y, int c0, int c1) {
  for (int64_t vox = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vox < x.vox; vox += (int64_t)gridDim.x * blockDim.x) {
    const TI* p = x.p + vox * x.ld;
    float mx = -INFINITY;
    for (int c = c0; c < c1; ++c) mx = fmaxf(mx, to_f<TI>(p[c]));
    float se = 0.f;
    for (int c = c0; c < c1; ++c) se += expf(to_f<TI>(p[c]) - mx);
    float inv = 1.f / se;
    TO* o = y.p + vox * y.ld;
    for (int c = c0; c < c1; ++c) o[c] = from_f<TO>(expf(to_f<TI>(p[c]) - mx) * inv);
  }
}

// ------------------------------------------------------------------------------------------------ optimiser
// adamw_kernel / adam_kernel / sgd_kernel take their hyper-parameters by value (one launch per host-side step);
// optim_prepare_kernel + optim_dev_kernel read them from device memory so that a captured CUDA graph can replay the update
// while the LR scheduler rewrites the block between replays (engine/train.py).
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt, float grad_scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);
    float mi = m[i] + (1.f - b1) * (gi - m[i]);          // torch: exp_avg.lerp_(grad, 1 - beta1)
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

// torch.optim.Adam (timm create_optimizer_v2('adam', weight_decay=wd), reference engine/__init__.py:58-70): the decay is an L2
// term added to the gradient before the moments; `decoupled` = 1 is AdamW's shrink of the parameter instead.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale,
                            int decoupled) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * grad_scale;
    if (decoupled) pi *= (1.f - lr * wd);
    else gi = fmaf(wd, pi, gi);
    const float mi = m[i] + (1.f - b1) * (gi - m[i]);
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

// torch.optim.SGD: g += wd p; buf = first ? g : momentum buf + g; g = nesterov ? g + momentum buf : buf; p -= lr g
// (timm's 'sgd' is SGD(momentum=0.9, nesterov=True)).
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, int64_t n,
                           float lr, float momentum, float wd, int first, float grad_scale, int nesterov) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale + wd * p[i];
    if (momentum != 0.f) {
      float b = first ? gi : momentum * mom[i] + gi;
      mom[i] = b;
      gi = nesterov ? fmaf(momentum, b, gi) : b;
    }
    p[i] -= lr * gi;
  }
}

// Device-resident hyper-parameters.  hp[B200_HP_*] (float): lr, beta1, beta2, eps, weight_decay, momentum, nesterov, grad_scale,
// clip_norm.  state (int64): [0] optimiser steps taken, [1] steps skipped because the gradient was not finite (fp16 overflow).
// derived (float): [0] 1 - beta1^t, [1] sqrt(1 - beta2^t), [2] grad_scale incl. clipping and 1 / *denom, [3] skip, [4] first step.
// gsq: sum of squares of the raw gradient (clip_grad_norm_, train_engine.py:174-176, and the fp16 overflow test) or null;
// denom: divisor of the gradient known only on the device (Noise2Void mask count) or null.
__global__ void optim_prepare_kernel(const float* __restrict__ hp, const double* __restrict__ gsq, const double* __restrict__ denom,
                                     int64_t* __restrict__ state, float* __restrict__ derived) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double scale = (double)hp[7];
  if (denom) scale /= *denom;
  int skip = !(scale == scale) || scale > 3.0e38 || scale < -3.0e38;
  if (gsq) {
    const double total = sqrt(*gsq) * fabs(scale);
    if (!(total == total) || total > 1.0e300) skip = 1;
    else if (hp[8] > 0.f) {
      const double coef = (double)hp[8] / (total + 1e-6);
      if (coef < 1.0) scale *= coef;
    }
  }
  if (skip) { state[1] += 1; derived[3] = 1.f; return; }
  const int64_t t = ++state[0];
  derived[0] = (float)(1.0 - pow((double)hp[1], (double)t));
  derived[1] = (float)sqrt(1.0 - pow((double)hp[2], (double)t));
  derived[2] = (float)scale;
  derived[3] = 0.f;
  derived[4] = t == 1 ? 1.f : 0.f;
}

template <int KIND>   // 0 AdamW, 1 Adam (L2), 2 SGD
__global__ void optim_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 int64_t n, const float* __restrict__ hp, const float* __restrict__ derived) {
  if (derived[3] != 0.f) return;
  const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4], momentum = hp[5];
  const int nesterov = hp[6] != 0.f, first = derived[4] != 0.f;
  const float bc1 = derived[0], bc2_sqrt = derived[1], grad_scale = derived[2];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * grad_scale;
    if (KIND == 2) {
      gi = fmaf(wd, pi, gi);
      if (momentum != 0.f) {
        const float b = first ? gi : momentum * m[i] + gi;
        m[i] = b;
        gi = nesterov ? fmaf(momentum, b, gi) : b;
      }
      p[i] = pi - lr * gi;
    } else {
      if (KIND == 0) pi *= (1.f - lr * wd);
      else gi = fmaf(wd, pi, gi);
      const float mi = m[i] + (1.f - b1) * (gi - m[i]);
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      pi -= (lr / bc1) * (mi / denom);
      p[i] = pi; m[i] = mi; v[i] = vi;
    }
  }
}

struct FloatBlock { float v[16]; };
__global__ void write_floats_kernel(float* __restrict__ dst, FloatBlock b, int n) {
  if (threadIdx.x < n) dst[threadIdx.x] = b.v[threadIdx.x];
}

__global__ void scale_by_dev_kernel(float* __restrict__ g, int64_t n, const double* __restrict__ denom, float mul) {
  const float f = (float)((double)mul / *denom);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g[i] *= f;
}

__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += (double)g[i] * g[i];
  block_atomic_add(acc, out);
}

}  // namespace b200

using namespace b200;

static inline unsigned rows_grid(int64_t spatial, int rows, int n) {
  int64_t b = ceil_div(spatial, (int64_t)rows * 8);
  int64_t cap = ceil_div((int64_t)sm_count() * 8, n);
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ================================================================================================ C ABI
B200_EXPORT int b200_channel_sums(const b200_tensor* x, double* sums, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "channel_sums.x") && sums, "%s", b200_last_error());
  B200_DISPATCH_DTYPE(x->dtype, T, return launch_channel_sums<T>(x, sums, (cudaStream_t)stream));
  return B200_OK;
}

B200_EXPORT int b200_norm_finalize(const double* sums, int32_t n, int32_t c, int32_t groups, int64_t spatial,
                                   int32_t batch_stats, const float* gamma, const float* beta, float eps,
                                   float* mean, float* rstd, float* scale, float* shift, void* stream) {
  B200_CHECK_ARG(sums && mean && rstd && scale && shift, "norm_finalize: null pointer");
  B200_CHECK_ARG(groups > 0 && c % groups == 0, "norm_finalize: channels %d not divisible by groups %d", c, groups);
  int total = n * groups * 32;                    // one warp per (n, g)
  norm_finalize_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, n, c, groups, spatial, batch_stats,
                                                                             gamma, beta, eps, mean, rstd, scale, shift);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_scale_shift_act(const b200_tensor* x, const float* scale, const float* shift, int32_t act,
                                     const b200_tensor* y, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "scale_shift_act.x") && check_tensor(y, "scale_shift_act.y"), "%s", b200_last_error());
  B200_CHECK_ARG(same_spatial(x, y) && x->c == y->c, "scale_shift_act: shape mismatch");
  B200_CHECK_ARG((scale == nullptr) == (shift == nullptr), "scale_shift_act: scale/shift must both be given");
  cudaStream_t st = (cudaStream_t)stream;
#define SSA(TI, TO)                                                                                              \
  {                                                                                                              \
    constexpr int V = VecOf<TI>::n < VecOf<TO>::n ? VecOf<TI>::n : VecOf<TO>::n;                                  \
    View<const TI> xv{(const TI*)x->data, x->ld, x->c, voxels(x), (int64_t)x->d * x->h * x->w};                    \
    View<TO> yv{(TO*)y->data, y->ld, y->c, voxels(y), (int64_t)y->d * y->h * y->w};                                \
    if (scale && std::is_same<TI, TO>::value && vec_ok(x, VecOf<TI>::n) && vec_ok(y, VecOf<TI>::n) &&             \
        x->c / VecOf<TI>::n <= 256) {                                                                             \
      constexpr int VV = VecOf<TI>::n;                                                                            \
      int cvn = x->c / VV, rows = 256 / cvn;                                                                      \
      dim3 grid(rows_grid(xv.spatial, rows, x->n), x->n);                                                         \
      View<TI> yv2{(TI*)y->data, y->ld, y->c, voxels(y), (int64_t)y->d * y->h * y->w};                              \
      B200_DISPATCH_ACT(act, ACT, (scale_shift_act_rows_kernel<TI, VV, ACT><<<grid, 256, 0, st>>>(xv, yv2, scale, shift, act, cvn, rows))); \
    } else if (vec_ok(x, V) && vec_ok(y, V) && sizeof(TI) == sizeof(TO))                                          \
      scale_shift_act_kernel<TI, TO, V><<<grid_for(xv.vox * (x->c / V), 256), 256, 0, st>>>(xv, yv, scale, shift, act); \
    else                                                                                                         \
      scale_shift_act_kernel<TI, TO, 1><<<grid_for(xv.vox * x->c, 256), 256, 0, st>>>(xv, yv, scale, shift, act);  \
  }
  B200_DISPATCH_DTYPE(x->dtype, TI_, { B200_DISPATCH_DTYPE(y->dtype, TO_, SSA(TI_, TO_)); });
#undef SSA
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_convert(const b200_tensor* src, const b200_tensor* dst, void* stream) {
  return b200_scale_shift_act(src, nullptr, nullptr, B200_ACT_NONE, dst, stream);
}

B200_EXPORT int b200_norm_act_bwd_reduce(const b200_tensor* x, const b200_tensor* dy, const float* mean,
                                         const float* rstd, int32_t groups, const float* gamma, const float* beta,
                                         int32_t act, double* red, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "bwd_reduce.x") && check_tensor(dy, "bwd_reduce.dy") && mean && rstd && red, "%s",
                 b200_last_error());
  B200_CHECK_ARG(same_spatial(x, dy) && x->c == dy->c && x->dtype == dy->dtype, "bwd_reduce: shape/dtype mismatch");
  B200_CHECK_ARG(groups > 0 && x->c % groups == 0, "bwd_reduce: bad groups");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t spatial = (int64_t)x->d * x->h * x->w;
  auto grid_of = [&](int rows) {
    int64_t chunks = ceil_div(spatial, (int64_t)rows * 8);
    int64_t cap = ceil_div((int64_t)sm_count() * 4, x->n);
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    return dim3((unsigned)chunks, x->n);
  };
  B200_DISPATCH_DTYPE(x->dtype, T, {
    constexpr int V = VecOf<T>::n;
    constexpr int VH = V;
    View<const T> xv{(const T*)x->data, x->ld, x->c, voxels(x), spatial};
    View<const T> dv{(const T*)dy->data, dy->ld, dy->c, voxels(dy), spatial};
    if (vec_ok(x, V) && vec_ok(dy, V) && x->c / VH <= 256) {
      int cvn = x->c / VH, rows = 256 / cvn;
      int threads = ((rows * cvn + 31) / 32) * 32;
      static const int variant = getenv("B200_ROWS_VARIANT") ? atoi(getenv("B200_ROWS_VARIANT")) : 0;
      const size_t smem = sizeof(double) * rows * cvn * VH * 2;
      B200_DISPATCH_ACT(act, ACT, {
        if (variant == 1)
          norm_act_bwd_reduce_kernel<T, VH, ACT, 2, 3><<<grid_of(rows), threads, smem, st>>>(xv, dv, mean, rstd, groups, gamma, beta, act, red, cvn, rows);
        else if (variant == 2)
          norm_act_bwd_reduce_kernel<T, VH, ACT, 2, 4><<<grid_of(rows), threads, smem, st>>>(xv, dv, mean, rstd, groups, gamma, beta, act, red, cvn, rows);
        else
          norm_act_bwd_reduce_kernel<T, VH, ACT, 1, 4><<<grid_of(rows), threads, smem, st>>>(xv, dv, mean, rstd, groups, gamma, beta, act, red, cvn, rows);
      });
    } else {
      dim3 grid = grid_of(1);
      B200_CHECK_ARG(x->c <= 1024, "bwd_reduce: too many channels");
      int cvn = x->c, rows = 256 / cvn;
      if (rows < 1) rows = 1;
      if (rows > 32) rows = 32;
      int threads = ((rows * cvn + 31) / 32) * 32;
      norm_act_bwd_reduce_kernel<T, 1, -1, 1, 4><<<grid, threads, sizeof(double) * rows * cvn * 2, st>>>(
          xv, dv, mean, rstd, groups, gamma, beta, act, red, cvn, rows);
    }
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_norm_bwd_finalize(const double* red, const float* mean, const float* rstd, const float* gamma,
                                       const float* beta, int32_t n, int32_t c, int32_t groups, int64_t spatial,
                                       int32_t batch_stats, float* coef, float* dgamma, float* dbeta, void* stream) {
  B200_CHECK_ARG(red && mean && rstd && coef, "norm_bwd_finalize: null pointer");
  int total = n * groups * 32 > c ? n * groups * 32 : c;      // one warp per (n, g), one thread per channel
  norm_bwd_finalize_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(red, mean, rstd, gamma, beta, n, c, groups,
                                                                                 spatial, batch_stats, coef, dgamma, dbeta);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_norm_bwd_finalize_sums(const double* red, const float* mean, const float* rstd, const float* gamma,
                                            const float* beta, int32_t n, int32_t c, int32_t groups, int64_t spatial,
                                            int32_t batch_stats, float* coef, float* dgamma, float* dbeta, const double* xsums,
                                            float* dxsum, void* stream) {
  B200_CHECK_ARG(red && mean && rstd && coef && xsums && dxsum, "norm_bwd_finalize_sums: null pointer");
  int total = n * groups * 32 > c ? n * groups * 32 : c;
  norm_bwd_finalize_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(red, mean, rstd, gamma, beta, n, c, groups,
                                                                                 spatial, batch_stats, coef, dgamma, dbeta, xsums, dxsum);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_sums_through_pointwise(const float* w, const float* dysum, float* xsum, int32_t cout, int32_t cin, void* stream) {
  B200_CHECK_ARG(w && dysum && xsum && cout > 0 && cin > 0, "sums_through_pointwise: bad args");
  sums_through_pointwise_kernel<<<(cin * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w, dysum, xsum, cout, cin);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------ dropout
// nn.Dropout(p) in training mode (reference blocks.py:162-163): y = keep ? x / (1 - p) : 0.  The keep decision is a pure
// function of (seed, layer id, logical element index), so the backward pass re-derives the mask instead of storing it and
// a replayed CUDA graph gets fresh masks by bumping the device-resident seed.  Counter-based generator: two rounds of
// the splitmix64 finaliser over (seed, layer, index) -- not torch's Philox stream, so masks differ from the reference's
// bit for bit (they are random numbers there too); the distribution and the scaling are the same.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t layer, uint64_t idx, uint32_t thresh) {
  const uint64_t h = mix64(mix64(seed + 0x9E3779B97F4A7C15ull * (layer + 1)) ^ (idx * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull));
  return (uint32_t)(h >> 32) >= thresh;              // P(keep) = 1 - p
}

template <typename T>
__global__ void dropout_kernel(View<const T> x, View<T> y, float scale, uint32_t thresh, const int64_t* __restrict__ seed_dev,
                               uint64_t layer, int accumulate) {
  const uint64_t seed = (uint64_t)*seed_dev;
  const int64_t total = x.vox * x.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t vox = i / x.c;
    const int c = (int)(i - vox * x.c);
    const float v = dropout_keep(seed, layer, (uint64_t)i, thresh) ? to_f<T>(x.p[vox * x.ld + c]) * scale : 0.f;
    T* o = y.p + vox * y.ld + c;
    *o = from_f<T>(accumulate ? to_f<T>(*o) + v : v);
  }
}

B200_EXPORT int b200_dropout(const b200_tensor* x, const b200_tensor* y, float p, const int64_t* seed_dev, int64_t layer_id,
                             int32_t accumulate, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "dropout.x") && check_tensor(y, "dropout.y") && seed_dev, "%s", b200_last_error());
  B200_CHECK_ARG(same_spatial(x, y) && x->c == y->c && x->dtype == y->dtype, "dropout: shape/dtype mismatch");
  B200_CHECK_ARG(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
  const double t = (double)p * 4294967296.0;
  const uint32_t thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  B200_DISPATCH_DTYPE(x->dtype, T, (dropout_kernel<T><<<grid_for(voxels(x) * x->c, 256), 256, 0, (cudaStream_t)stream>>>(
                                       view<const T>(x), view<T>(y), 1.f / (1.f - p), thresh, seed_dev, (uint64_t)layer_id, accumulate)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ------------------------------------------------------------------------------ linear up-sampling (nn.Upsample)
// mode = 'bilinear' / 'trilinear', align_corners = False, integer scale factors (reference blocks.py:604-606):
// src = max((dst + 0.5) / s - 0.5, 0), i0 = floor(src), i1 = min(i0 + 1, in - 1), y = (1 - l) * x[i0] + l * x[i1] per axis.
struct LinAxis { int i0, i1; float l; };
__device__ __forceinline__ LinAxis lin_axis(int o, int s, int in) {
  float src = ((float)o + 0.5f) / (float)s - 0.5f;
  if (src < 0.f) src = 0.f;
  int i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  LinAxis a;
  a.i0 = i0;
  a.i1 = i0 < in - 1 ? i0 + 1 : i0;
  a.l = src - (float)i0;
  return a;
}
// weight of input index i in output index o along one axis
__device__ __forceinline__ float lin_weight(int o, int i, int s, int in) {
  const LinAxis a = lin_axis(o, s, in);
  return (a.i0 == i ? 1.f - a.l : 0.f) + (a.i1 == i ? a.l : 0.f);
}

template <typename T>
__global__ void upsample_linear_fwd_kernel(const T* __restrict__ x, int64_t ldx, T* __restrict__ y, int64_t ldy, int n, int d, int h,
                                           int w, int c, int sd, int sh, int sw) {
  const int od = d * sd, oh = h * sh, ow = w * sw;
  const int64_t total = (int64_t)n * od * oh * ow * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int ch = (int)(t % c); t /= c;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh); t /= oh;
    const int oz = (int)(t % od);
    const int nn = (int)(t / od);
    const LinAxis az = lin_axis(oz, sd, d), ay = lin_axis(oy, sh, h), ax = lin_axis(ox, sw, w);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int iz = a ? az.i1 : az.i0, iy = b ? ay.i1 : ay.i0, ix = e ? ax.i1 : ax.i0;
          const float wt = (a ? az.l : 1.f - az.l) * (b ? ay.l : 1.f - ay.l) * (e ? ax.l : 1.f - ax.l);
          acc = fmaf(wt, to_f<T>(x[((((int64_t)nn * d + iz) * h + iy) * w + ix) * ldx + ch]), acc);
        }
    y[((((int64_t)nn * od + oz) * oh + oy) * ow + ox) * ldy + ch] = from_f<T>(acc);
  }
}

// gather form of the adjoint: every input element sums the output gradients it contributed to
template <typename T>
__global__ void upsample_linear_bwd_kernel(const T* __restrict__ dy, int64_t lddy, T* __restrict__ dx, int64_t lddx, int n, int d,
                                           int h, int w, int c, int sd, int sh, int sw, int accumulate) {
  const int od = d * sd, oh = h * sh, ow = w * sw;
  const int64_t total = (int64_t)n * d * h * w * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int ch = (int)(t % c); t /= c;
    const int ix = (int)(t % w); t /= w;
    const int iy = (int)(t % h); t /= h;
    const int iz = (int)(t % d);
    const int nn = (int)(t / d);
    const int z0 = max(0, sd * iz - sd), z1 = min(od, sd * iz + 2 * sd);
    const int y0 = max(0, sh * iy - sh), y1 = min(oh, sh * iy + 2 * sh);
    const int x0 = max(0, sw * ix - sw), x1 = min(ow, sw * ix + 2 * sw);
    float acc = 0.f;
    for (int oz = z0; oz < z1; ++oz) {
      const float wz = lin_weight(oz, iz, sd, d);
      if (wz == 0.f) continue;
      for (int oy = y0; oy < y1; ++oy) {
        const float wy = lin_weight(oy, iy, sh, h);
        if (wy == 0.f) continue;
        for (int ox = x0; ox < x1; ++ox) {
          const float wx = lin_weight(ox, ix, sw, w);
          if (wx == 0.f) continue;
          acc = fmaf(wz * wy * wx, to_f<T>(dy[((((int64_t)nn * od + oz) * oh + oy) * ow + ox) * lddy + ch]), acc);
        }
      }
    }
    T* o = dx + ((((int64_t)nn * d + iz) * h + iy) * w + ix) * lddx + ch;
    *o = from_f<T>(accumulate ? to_f<T>(*o) + acc : acc);
  }
}

static int upsample_geom(const b200_tensor* coarse, const b200_tensor* fine, int* sd, int* sh, int* sw) {
  B200_CHECK_ARG(coarse->n == fine->n && coarse->c == fine->c && coarse->dtype == fine->dtype &&
                     fine->d % coarse->d == 0 && fine->h % coarse->h == 0 && fine->w % coarse->w == 0,
                 "upsample_linear: fine dims must be integer multiples of the coarse dims, same N / C / dtype");
  *sd = fine->d / coarse->d; *sh = fine->h / coarse->h; *sw = fine->w / coarse->w;
  return B200_OK;
}

B200_EXPORT int b200_upsample_linear_fwd(const b200_tensor* x, const b200_tensor* y, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "upsample.x") && check_tensor(y, "upsample.y"), "%s", b200_last_error());
  int sd, sh, sw;
  int rc = upsample_geom(x, y, &sd, &sh, &sw);
  if (rc) return rc;
  B200_DISPATCH_DTYPE(x->dtype, T, (upsample_linear_fwd_kernel<T><<<grid_for(voxels(y) * y->c, 256), 256, 0, (cudaStream_t)stream>>>(
                                       (const T*)x->data, x->ld, (T*)y->data, y->ld, x->n, x->d, x->h, x->w, x->c, sd, sh, sw)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_upsample_linear_bwd(const b200_tensor* dy, const b200_tensor* dx, int32_t accumulate, void* stream) {
  B200_CHECK_ARG(check_tensor(dy, "upsample_bwd.dy") && check_tensor(dx, "upsample_bwd.dx"), "%s", b200_last_error());
  int sd, sh, sw;
  int rc = upsample_geom(dx, dy, &sd, &sh, &sw);
  if (rc) return rc;
  B200_DISPATCH_DTYPE(dx->dtype, T, (upsample_linear_bwd_kernel<T><<<grid_for(voxels(dx) * dx->c, 256), 256, 0, (cudaStream_t)stream>>>(
                                        (const T*)dy->data, dy->ld, (T*)dx->data, dx->ld, dx->n, dx->d, dx->h, dx->w, dx->c, sd, sh,
                                        sw, accumulate)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// BatchNorm bookkeeping (tiny): running statistics from the batch statistics of b200_norm_finalize(batch_stats = 1), and
// the per-(n, c) scale / shift of eval mode from the running statistics
__global__ void bn_update_running_kernel(const float* __restrict__ mean, const float* __restrict__ rstd, float eps, double count,
                                         float momentum, float* __restrict__ rm, float* __restrict__ rv, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double r = (double)rstd[i];
  double var = 1.0 / (r * r) - (double)eps;                 // biased batch variance
  if (var < 0.0) var = 0.0;
  const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
  rm[i] = (float)((1.0 - (double)momentum) * (double)rm[i] + (double)momentum * (double)mean[i]);
  rv[i] = (float)((1.0 - (double)momentum) * (double)rv[i] + (double)momentum * unbiased);
}

__global__ void bn_eval_coeffs_kernel(const float* __restrict__ rm, const float* __restrict__ rv, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float eps, int n, int c, float* __restrict__ scale,
                                      float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int cc = i % c;
  const float sc = (gamma ? gamma[cc] : 1.f) / sqrtf(rv[cc] + eps);
  scale[i] = sc;
  shift[i] = (beta ? beta[cc] : 0.f) - rm[cc] * sc;
}

B200_EXPORT int b200_bn_update_running(const float* mean, const float* rstd, float eps, double count, float momentum,
                                       float* running_mean, float* running_var, int32_t c, void* stream) {
  B200_CHECK_ARG(mean && rstd && running_mean && running_var && c > 0 && count > 0, "bn_update_running: bad args");
  bn_update_running_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean, rstd, eps, count, momentum, running_mean,
                                                                             running_var, c);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_bn_eval_coeffs(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                                    float eps, int32_t n, int32_t c, float* scale, float* shift, void* stream) {
  B200_CHECK_ARG(running_mean && running_var && scale && shift && n > 0 && c > 0, "bn_eval_coeffs: bad args");
  bn_eval_coeffs_kernel<<<(n * c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, gamma, beta, eps, n, c,
                                                                              scale, shift);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_norm_act_bwd_apply(const b200_tensor* x, const b200_tensor* dy, int32_t act, const float* coef,
                                        const b200_tensor* dx, int32_t accumulate, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "bwd_apply.x") && check_tensor(dy, "bwd_apply.dy") && check_tensor(dx, "bwd_apply.dx") && coef,
                 "%s", b200_last_error());
  B200_CHECK_ARG(same_spatial(x, dy) && same_spatial(x, dx) && x->c == dy->c && x->c == dx->c &&
                     x->dtype == dy->dtype && x->dtype == dx->dtype, "bwd_apply: shape/dtype mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  B200_DISPATCH_DTYPE(x->dtype, T, {
    constexpr int V = VecOf<T>::n;
    View<const T> xv = view<const T>(x);
    View<const T> dv = view<const T>(dy);
    View<T> ov = view<T>(dx);
    if (vec_ok(x, V) && vec_ok(dy, V) && vec_ok(dx, V) && x->c / V <= 256) {
      int cvn = x->c / V, rows = 256 / cvn;
      dim3 grid(rows_grid(xv.spatial, rows, x->n), x->n);
      static const int variant = getenv("B200_ROWS_VARIANT") ? atoi(getenv("B200_ROWS_VARIANT")) : 0;
      B200_DISPATCH_ACT(act, ACT, {
        // measured (profiles/rows_micro.py, c48 @128^3 x4): U=2/2 blocks 3.5 TB/s, U=1/3 blocks 4.5, U=1/4 blocks 5.2 -- occupancy wins
        if (variant == 1) norm_act_bwd_apply_rows_kernel<T, V, ACT, 2, 4><<<grid, 256, 0, st>>>(xv, dv, ov, act, coef, accumulate, cvn, rows);
        else norm_act_bwd_apply_rows_kernel<T, V, ACT, 1, 4><<<grid, 256, 0, st>>>(xv, dv, ov, act, coef, accumulate, cvn, rows);
      });
    } else {
      norm_act_bwd_apply_kernel<T><<<grid_for(xv.vox * x->c, 256), 256, 0, st>>>(xv, dv, ov, act, coef, accumulate);
    }
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ------------------------------------------------------------------------------------- fast SiLU norm chain (16-bit dtypes)
// One-MUFU sigmoid and a backward that evaluates the activation derivative once (norm_fast.cuh).  Explicit entry points because
// the contract differs from b200_norm_act_bwd_*: the reduce pass OVERWRITES dy with g = dy * silu'(norm(x)) and the apply pass
// must be the matching one.  b200_norm_silu_fast_ok tells the caller whether the three tensors qualify.
static bool norm_fast_ok(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx) {
  if (!x || x->dtype == B200_F32) return false;
  const int V = 8;
  if (!vec_ok(x, V) || x->c / V > 256) return false;
  if (dy && !(vec_ok(dy, V) && dy->dtype == x->dtype && dy->c == x->c && same_spatial(x, dy))) return false;
  if (dx && !(vec_ok(dx, V) && dx->dtype == x->dtype && dx->c == x->c && same_spatial(x, dx))) return false;
  return true;
}

B200_EXPORT int b200_norm_silu_fast_ok(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx) {
  return norm_fast_ok(x, dy, dx) ? 1 : 0;
}

B200_EXPORT int b200_scale_shift_silu_fast(const b200_tensor* x, const float* scale, const float* shift, const b200_tensor* y,
                                           void* stream) {
  B200_CHECK_ARG(check_tensor(x, "silu_fast.x") && check_tensor(y, "silu_fast.y") && scale && shift, "%s", b200_last_error());
  B200_CHECK_ARG(norm_fast_ok(x, y, nullptr), "scale_shift_silu_fast: tensors do not qualify (16-bit, 8-channel vectors)");
  cudaStream_t st = (cudaStream_t)stream;
  B200_DISPATCH_DTYPE16(x->dtype, T, {
    constexpr int V = 8;
    View<const T> xv = view<const T>(x);
    View<T> yv = view<T>(y);
    int cvn = x->c / V, rows = 256 / cvn;
    dim3 grid(rows_grid(xv.spatial, rows, x->n), x->n);
    scale_shift_silu_rows_fast_kernel<T, V><<<grid, 256, 0, st>>>(xv, yv, scale, shift, cvn, rows);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_norm_silu_bwd_reduce_g(const b200_tensor* x, const b200_tensor* dy_g, const float* mean, const float* rstd,
                                            int32_t groups, const float* gamma, const float* beta, double* red, int32_t write_g,
                                            void* stream) {
  B200_CHECK_ARG(check_tensor(x, "reduce_g.x") && check_tensor(dy_g, "reduce_g.dy") && mean && rstd && red, "%s", b200_last_error());
  B200_CHECK_ARG(norm_fast_ok(x, dy_g, nullptr), "norm_silu_bwd_reduce_g: tensors do not qualify");
  B200_CHECK_ARG(groups > 0 && x->c % groups == 0, "norm_silu_bwd_reduce_g: bad groups");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t spatial = (int64_t)x->d * x->h * x->w;
  B200_DISPATCH_DTYPE16(x->dtype, T, {
    constexpr int V = 8;
    View<const T> xv{(const T*)x->data, x->ld, x->c, voxels(x), spatial};
    View<T> gv{(T*)dy_g->data, dy_g->ld, dy_g->c, voxels(dy_g), spatial};
    int cvn = x->c / V, rows = 256 / cvn;
    int threads = ((rows * cvn + 31) / 32) * 32;
    int64_t chunks = ceil_div(spatial, (int64_t)rows * 8);
    int64_t cap = ceil_div((int64_t)sm_count() * 4, x->n);
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    const size_t smem = sizeof(double) * rows * cvn * V * 2;
    if (write_g)
      norm_silu_bwd_reduce_g_kernel<T, V, 1, 4, true><<<dim3((unsigned)chunks, x->n), threads, smem, st>>>(xv, gv, mean, rstd, groups,
                                                                                                       gamma, beta, red, cvn, rows);
    else
      norm_silu_bwd_reduce_g_kernel<T, V, 1, 4, false><<<dim3((unsigned)chunks, x->n), threads, smem, st>>>(xv, gv, mean, rstd, groups,
                                                                                                        gamma, beta, red, cvn, rows);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_norm_silu_bwd_apply_fast(const b200_tensor* x, const b200_tensor* dy, const float* coef, const b200_tensor* dx,
                                              int32_t accumulate, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "apply_fast.x") && check_tensor(dy, "apply_fast.dy") && check_tensor(dx, "apply_fast.dx") && coef, "%s",
                 b200_last_error());
  B200_CHECK_ARG(norm_fast_ok(x, dy, dx), "norm_silu_bwd_apply_fast: tensors do not qualify");
  cudaStream_t st = (cudaStream_t)stream;
  B200_DISPATCH_DTYPE16(x->dtype, T, {
    constexpr int V = 8;
    View<const T> xv = view<const T>(x);
    View<const T> dv = view<const T>(dy);
    View<T> ov = view<T>(dx);
    int cvn = x->c / V, rows = 256 / cvn;
    dim3 grid(rows_grid(xv.spatial, rows, x->n), x->n);
    norm_silu_bwd_apply_fast_rows_kernel<T, V, 1, 4><<<grid, 256, 0, st>>>(xv, dv, ov, coef, accumulate, cvn, rows);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_norm_bwd_apply_g(const b200_tensor* x, const b200_tensor* g, const float* coef, const b200_tensor* dx,
                                      int32_t accumulate, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "apply_g.x") && check_tensor(g, "apply_g.g") && check_tensor(dx, "apply_g.dx") && coef, "%s",
                 b200_last_error());
  B200_CHECK_ARG(norm_fast_ok(x, g, dx), "norm_bwd_apply_g: tensors do not qualify");
  cudaStream_t st = (cudaStream_t)stream;
  B200_DISPATCH_DTYPE16(x->dtype, T, {
    constexpr int V = 8;
    View<const T> xv = view<const T>(x);
    View<const T> gv = view<const T>(g);
    View<T> ov = view<T>(dx);
    int cvn = x->c / V, rows = 256 / cvn;
    dim3 grid(rows_grid(xv.spatial, rows, x->n), x->n);
    norm_bwd_apply_g_rows_kernel<T, V, 1, 4><<<grid, 256, 0, st>>>(xv, gv, ov, coef, accumulate, cvn, rows);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_act_bwd(const b200_tensor* x, const b200_tensor* dy, int32_t act, const b200_tensor* dx,
                             int32_t accumulate, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "act_bwd.x") && check_tensor(dy, "act_bwd.dy") && check_tensor(dx, "act_bwd.dx"), "%s",
                 b200_last_error());
  B200_CHECK_ARG(same_spatial(x, dy) && same_spatial(x, dx) && x->c == dy->c && x->c == dx->c &&
                     x->dtype == dy->dtype && x->dtype == dx->dtype, "act_bwd: shape/dtype mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  B200_DISPATCH_DTYPE(x->dtype, T, {
    constexpr int V = VecOf<T>::n;
    View<const T> xv = view<const T>(x);
    View<const T> dv = view<const T>(dy);
    View<T> ov = view<T>(dx);
    if (vec_ok(x, V) && vec_ok(dy, V) && vec_ok(dx, V))
      act_bwd_kernel<T, V><<<grid_for(xv.vox * (x->c / V), 256), 256, 0, st>>>(xv, dv, ov, act, accumulate);
    else
      act_bwd_kernel<T, 1><<<grid_for(xv.vox * x->c, 256), 256, 0, st>>>(xv, dv, ov, act, accumulate);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static int pool_geom(const b200_tensor* x, const b200_tensor* y, int pd, int ph, int pw, PoolGeom* g) {
  B200_CHECK_ARG(pd > 0 && ph > 0 && pw > 0, "maxpool: bad window");
  B200_CHECK_ARG(y->n == x->n && y->c == x->c && y->d == x->d / pd && y->h == x->h / ph && y->w == x->w / pw,
                 "maxpool: output shape (%d,%d,%d,%d,%d) does not match floor(input/window)", y->n, y->d, y->h, y->w, y->c);
  *g = PoolGeom{x->n, x->d, x->h, x->w, x->c, y->d, y->h, y->w, pd, ph, pw};
  return B200_OK;
}

B200_EXPORT int b200_maxpool_fwd(const b200_tensor* x, const b200_tensor* y, int32_t pd, int32_t ph, int32_t pw,
                                 void* stream) {
  B200_CHECK_ARG(check_tensor(x, "maxpool.x") && check_tensor(y, "maxpool.y") && x->dtype == y->dtype, "%s",
                 b200_last_error());
  PoolGeom g;
  int st = pool_geom(x, y, pd, ph, pw, &g);
  if (st) return st;
  int64_t total = voxels(y) * y->c;
  B200_DISPATCH_DTYPE(x->dtype, T, {
    constexpr int V = VecOf<T>::n;
    if (vec_ok(x, V) && vec_ok(y, V) && pool_win_enabled() && ph == 2 && pw == 2 && (pd == 1 || pd == 2)) {
      launch_pool_fwd_win<T, V>(x, y, g, pd, (cudaStream_t)stream);
    } else if (vec_ok(x, V) && vec_ok(y, V))
      maxpool_fwd_vec_kernel<T, V><<<grid_for(total / V, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x->data, x->ld, (T*)y->data,
                                                                                             y->ld, g);
    else
      maxpool_fwd_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x->data, x->ld, (T*)y->data, y->ld, g);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_maxpool_bwd(const b200_tensor* x, const b200_tensor* y, const b200_tensor* dy, const b200_tensor* dx,
                                 int32_t pd, int32_t ph, int32_t pw, int32_t accumulate, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "maxpool_bwd.x") && check_tensor(dy, "maxpool_bwd.dy") && check_tensor(dx, "maxpool_bwd.dx"),
                 "%s", b200_last_error());
  B200_CHECK_ARG(x->dtype == dy->dtype && x->dtype == dx->dtype && same_spatial(x, dx) && x->c == dx->c,
                 "maxpool_bwd: shape/dtype mismatch");
  PoolGeom g;
  int st = pool_geom(x, dy, pd, ph, pw, &g);
  if (st) return st;
  (void)y;
  int64_t total = voxels(x) * x->c;
  const bool divisible = x->d % pd == 0 && x->h % ph == 0 && x->w % pw == 0;
  B200_DISPATCH_DTYPE(x->dtype, T, {
    constexpr int V = VecOf<T>::n;
    if (divisible && vec_ok(x, V) && vec_ok(dy, V) && vec_ok(dx, V) && pool_win_enabled() && ph == 2 && pw == 2 &&
        (pd == 1 || pd == 2)) {
      launch_pool_bwd_win<T, V>(x, dy, dx, nullptr, g, pd, accumulate, (cudaStream_t)stream);
    } else if (divisible && vec_ok(x, V) && vec_ok(dy, V) && vec_ok(dx, V))
      maxpool_bwd_vec_kernel<T, V, 8><<<grid_for(voxels(dy) * (x->c / V), 256), 256, 0, (cudaStream_t)stream>>>(
          (const T*)x->data, x->ld, (const T*)dy->data, dy->ld, (T*)dx->data, dx->ld, g, accumulate);
    else
      maxpool_bwd_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x->data, x->ld, (const T*)dy->data,
                                                                                  dy->ld, (T*)dx->data, dx->ld, g, accumulate);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static bool pool_bwd_to_ok(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx_in, const b200_tensor* dx_out, int pd, int ph,
                           int pw) {
  if (!x || !dy || !dx_out || x->dtype == B200_F32 || !pool_win_enabled() || ph != 2 || pw != 2 || (pd != 1 && pd != 2)) return false;
  if (x->d % pd || x->h % ph || x->w % pw) return false;
  if (!(vec_ok(x, 8) && vec_ok(dy, 8) && vec_ok(dx_out, 8))) return false;
  if (dx_in && !(vec_ok(dx_in, 8) && dx_in->dtype == x->dtype && dx_in->c == x->c && same_spatial(x, dx_in))) return false;
  return dy->dtype == x->dtype && dx_out->dtype == x->dtype && dx_out->c == x->c && same_spatial(x, dx_out);
}

B200_EXPORT int b200_maxpool_bwd_to_ok(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx_in, const b200_tensor* dx_out,
                                       int32_t pd, int32_t ph, int32_t pw) {
  return pool_bwd_to_ok(x, dy, dx_in, dx_out, pd, ph, pw) ? 1 : 0;
}

B200_EXPORT int b200_maxpool_bwd_to(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx_in, const b200_tensor* dx_out,
                                    int32_t pd, int32_t ph, int32_t pw, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "maxpool_bwd_to.x") && check_tensor(dy, "maxpool_bwd_to.dy") && check_tensor(dx_out, "maxpool_bwd_to.dx_out"),
                 "%s", b200_last_error());
  B200_CHECK_ARG(pool_bwd_to_ok(x, dy, dx_in, dx_out, pd, ph, pw), "maxpool_bwd_to: operands not supported (query b200_maxpool_bwd_to_ok)");
  PoolGeom g;
  int st = pool_geom(x, dy, pd, ph, pw, &g);
  if (st) return st;
  B200_DISPATCH_DTYPE16(x->dtype, T, (launch_pool_bwd_win<T, 8>(x, dy, dx_out, dx_in, g, pd, dx_in ? 1 : 0, (cudaStream_t)stream)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_binary(const b200_tensor* a, const b200_tensor* b, const b200_tensor* y, int32_t op, void* stream) {
  B200_CHECK_ARG(check_tensor(a, "binary.a") && check_tensor(y, "binary.y"), "%s", b200_last_error());
  B200_CHECK_ARG(op >= 0 && op <= 4, "binary: bad op");
  const bool needs_b = (op != 3 && op != 4);
  if (needs_b) B200_CHECK_ARG(check_tensor(b, "binary.b") && same_spatial(a, b) && (b->c == a->c || b->c == 1) &&
                                  b->dtype == a->dtype, "binary: operand mismatch");
  B200_CHECK_ARG(same_spatial(a, y) && a->c == y->c && a->dtype == y->dtype, "binary: output mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  const b200_tensor* bb = needs_b ? b : a;
  int bcast = needs_b && b->c == 1 && a->c != 1;
  if (a->dtype != B200_F32 && !bcast && vec_ok(a, 8) && vec_ok(bb, 8) && vec_ok(y, 8) && op >= 0 && op <= 4) {
    int64_t total = voxels(a) * (a->c / 8);
    int64_t blocks = ceil_div(total, 256 * 2);
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    if (blocks < 1) blocks = 1;
#define B200_BIN(T, OP) binary_vec_kernel<T, 8, OP><<<(unsigned)blocks, 256, 0, st>>>(view<const T>(a), view<const T>(bb), view<T>(y))
#define B200_BIN_OPS(T)                                                                                            \
    switch (op) { case 0: B200_BIN(T, 0); break; case 1: B200_BIN(T, 1); break; case 2: B200_BIN(T, 2); break;      \
                  case 3: B200_BIN(T, 3); break; default: B200_BIN(T, 4); break; }
    if (a->dtype == B200_BF16) { B200_BIN_OPS(__nv_bfloat16) } else { B200_BIN_OPS(__half) }
#undef B200_BIN_OPS
#undef B200_BIN
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  B200_DISPATCH_DTYPE(a->dtype, T, (binary_kernel<T, T, T><<<grid_for(voxels(a) * a->c, 256), 256, 0, st>>>(
                                       view<const T>(a), view<const T>(bb), view<T>(y), op, bcast)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_gate_bwd(const b200_tensor* x, const b200_tensor* psi, const b200_tensor* dout,
                              const b200_tensor* dpsi, const b200_tensor* dx, int32_t accumulate, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "gate.x") && check_tensor(psi, "gate.psi") && check_tensor(dout, "gate.dout") &&
                     check_tensor(dpsi, "gate.dpsi") && check_tensor(dx, "gate.dx"), "%s", b200_last_error());
  B200_CHECK_ARG(psi->c == 1 && dpsi->c == 1 && x->c == dout->c && x->c == dx->c && same_spatial(x, psi) &&
                     same_spatial(x, dout) && same_spatial(x, dpsi) && same_spatial(x, dx), "gate_bwd: shape mismatch");
  B200_DISPATCH_DTYPE(x->dtype, T, (gate_bwd_kernel<T><<<grid_for(voxels(x), 128), 128, 0, (cudaStream_t)stream>>>(
                                       view<const T>(x), view<const T>(psi), view<const T>(dout), view<T>(dpsi), view<T>(dx),
                                       accumulate)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_relu_mask_bwd(const b200_tensor* y, const b200_tensor* dy, const b200_tensor* da, void* stream) {
  B200_CHECK_ARG(check_tensor(y, "relu_mask.y") && check_tensor(dy, "relu_mask.dy") && check_tensor(da, "relu_mask.da"),
                 "%s", b200_last_error());
  B200_CHECK_ARG(same_spatial(y, dy) && same_spatial(y, da) && y->c == dy->c && y->c == da->c, "relu_mask_bwd: mismatch");
  B200_DISPATCH_DTYPE(y->dtype, T, (relu_mask_bwd_kernel<T><<<grid_for(voxels(y) * y->c, 256), 256, 0, (cudaStream_t)stream>>>(
                                       view<const T>(y), view<const T>(dy), view<T>(da))));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_bce_logits(const b200_tensor* logits, const float* target, double* loss_sum,
                                const b200_tensor* dlogits, float grad_scale, void* stream) {
  B200_CHECK_ARG(check_tensor(logits, "bce.logits") && target && loss_sum, "%s", b200_last_error());
  if (dlogits) B200_CHECK_ARG(check_tensor(dlogits, "bce.dlogits") && same_spatial(logits, dlogits) &&
                                  logits->c == dlogits->c && logits->dtype == dlogits->dtype, "bce: dlogits mismatch");
  const int64_t total = voxels(logits) * logits->c;
  if (logits->dtype != B200_F32 && logits->ld == logits->c && (!dlogits || dlogits->ld == dlogits->c) && total % 8 == 0 &&
      (((uintptr_t)logits->data | (uintptr_t)target | (uintptr_t)(dlogits ? dlogits->data : nullptr)) & 15) == 0) {
    B200_DISPATCH_DTYPE16(logits->dtype, T, {
      bce_logits_dense_kernel<T><<<grid_for(total / 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(
          (const T*)logits->data, target, loss_sum, dlogits ? (T*)dlogits->data : nullptr, grad_scale, total / 8);
    });
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  B200_DISPATCH_DTYPE(logits->dtype, T, {
    View<T> dv{dlogits ? (T*)dlogits->data : nullptr, dlogits ? dlogits->ld : 0, logits->c, voxels(logits), 0};
    bce_logits_kernel<T><<<grid_for(voxels(logits) * logits->c, 256, 4), 256, 0, (cudaStream_t)stream>>>(
        view<const T>(logits), target, loss_sum, dv, grad_scale);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_n2v_mse(const b200_tensor* pred, const float* target, double* sums, const b200_tensor* dpred,
                             float grad_scale, int32_t mode, void* stream) {
  B200_CHECK_ARG(check_tensor(pred, "n2v.pred") && target, "%s", b200_last_error());
  B200_CHECK_ARG(mode >= 0 && mode <= 2, "n2v: mode %d (0 sums, 1 gradient, 2 both)", mode);
  B200_CHECK_ARG((mode == 1 || sums) && (mode == 0 || (dpred && check_tensor(dpred, "n2v.dpred"))), "n2v: missing output");
  B200_DISPATCH_DTYPE(pred->dtype, T, {
    View<T> dv{dpred ? (T*)dpred->data : nullptr, dpred ? dpred->ld : 0, pred->c, voxels(pred), 0};
    n2v_mse_kernel<T><<<grid_for(voxels(pred) * pred->c, 256, 4), 256, 0, (cudaStream_t)stream>>>(
        view<const T>(pred), target, sums, dv, grad_scale, mode);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_softmax_ce(const b200_tensor* logits, const int64_t* target, double* sums,
                                const b200_tensor* dlogits, float grad_scale, int64_t ignore_index, void* stream) {
  B200_CHECK_ARG(check_tensor(logits, "ce.logits") && target && sums, "%s", b200_last_error());
  if (dlogits) B200_CHECK_ARG(check_tensor(dlogits, "ce.dlogits") && same_spatial(logits, dlogits) &&
                                  logits->c == dlogits->c && logits->dtype == dlogits->dtype, "ce: dlogits mismatch");
  B200_DISPATCH_DTYPE(logits->dtype, T, {
    View<T> dv{dlogits ? (T*)dlogits->data : nullptr, dlogits ? dlogits->ld : 0, logits->c, voxels(logits), 0};
    softmax_ce_kernel<T><<<grid_for(voxels(logits), 128, 4), 128, 0, (cudaStream_t)stream>>>(view<const T>(logits), target,
                                                                                          sums, dv, grad_scale, ignore_index);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_softmax_channels(const b200_tensor* x, const b200_tensor* y, int32_t c0, int32_t c1, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "softmax.x") && check_tensor(y, "softmax.y") && same_spatial(x, y) && x->c == y->c &&
                     x->dtype == y->dtype, "%s", b200_last_error());
  B200_CHECK_ARG(0 <= c0 && c0 < c1 && c1 <= x->c, "softmax: bad channel range");
  B200_DISPATCH_DTYPE(x->dtype, T, (softmax_channels_kernel<T, T><<<grid_for(voxels(x), 128), 128, 0, (cudaStream_t)stream>>>(
                                       view<const T>(x), view<T>(y), c0, c1)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
  B200_CHECK_ARG(p && g && m && v && n > 0 && step > 0, "adamw: bad args");
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                     bc1, bc2, grad_scale);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                               float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
  B200_CHECK_ARG(p && g && m && v && n > 0 && step > 0, "adam: bad args");
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                                                                    grad_scale, 0);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_sgd_step(float* p, const float* g, float* mom, int64_t n, float lr, float momentum,
                              float weight_decay, int32_t first_step, float grad_scale, int32_t nesterov, void* stream) {
  B200_CHECK_ARG(p && g && n > 0 && (momentum == 0.f || mom), "sgd: bad args");
  B200_CHECK_ARG(!nesterov || momentum > 0.f, "sgd: nesterov needs a momentum");
  sgd_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(p, g, mom, n, lr, momentum, weight_decay, first_step,
                                                                   grad_scale, nesterov);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_optim_step_dev(int32_t kind, float* p, const float* g, float* m, float* v, int64_t n, const float* hp,
                                    const double* gsq, const double* denom, int64_t* state, float* derived, void* stream) {
  B200_CHECK_ARG(p && g && hp && state && derived && n > 0, "optim_step_dev: null pointer");
  B200_CHECK_ARG(kind >= 0 && kind <= 2, "optim_step_dev: kind %d (0 AdamW, 1 Adam, 2 SGD)", kind);
  B200_CHECK_ARG(kind == 2 || (m && v), "optim_step_dev: Adam needs both moment buffers");
  cudaStream_t st = (cudaStream_t)stream;
  optim_prepare_kernel<<<1, 32, 0, st>>>(hp, gsq, denom, state, derived);
  if (kind == 0) optim_dev_kernel<0><<<grid_for(n, 256, 4), 256, 0, st>>>(p, g, m, v, n, hp, derived);
  else if (kind == 1) optim_dev_kernel<1><<<grid_for(n, 256, 4), 256, 0, st>>>(p, g, m, v, n, hp, derived);
  else optim_dev_kernel<2><<<grid_for(n, 256, 4), 256, 0, st>>>(p, g, m, v, n, hp, derived);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_pack_batch(const b200_pack_job* jobs, int32_t n_jobs, int32_t dtype, void* stream) {
  B200_CHECK_ARG(jobs && n_jobs > 0, "pack_batch: no jobs");
  B200_CHECK_ARG(dtype == B200_BF16 || dtype == B200_F16, "pack_batch: 16-bit engine dtypes only");
  static_assert(sizeof(b200_pack_job) == sizeof(PackJob), "b200_pack_job must mirror PackJob");
  cudaStream_t st = (cudaStream_t)stream;
  for (int j0 = 0; j0 < n_jobs; j0 += kPackBatchMax) {
    const int n = n_jobs - j0 < kPackBatchMax ? n_jobs - j0 : kPackBatchMax;
    PackJobTable tab;
    int block = 0;
    for (int k = 0; k < n; ++k) {
      const b200_pack_job& s = jobs[j0 + k];
      B200_CHECK_ARG(s.src && s.dst && s.kind >= 0 && s.kind <= 5 && s.cout > 0 && s.cin > 0 && s.kd > 0 && s.kh > 0 && s.kw > 0,
                     "pack_batch: bad job %d", j0 + k);
      PackJob& d = tab.j[k];
      d.src = s.src; d.dst = s.dst; d.kind = s.kind; d.cout = s.cout; d.cin = s.cin; d.kd = s.kd; d.kh = s.kh; d.kw = s.kw;
      d.flip = s.flip;
      int64_t total = (int64_t)s.cout * s.cin * s.kd * s.kh * s.kw;
      if (s.kind == PACK_XFOLD) {
        const int CO = s.flip ? s.cin : s.cout, CI = s.flip ? s.cout : s.cin;
        int xoff = 0, kxp = 0;
        B200_CHECK_ARG(pack_xfold_geom(CI, s.kw, &xoff, &kxp), "pack_batch: job %d: x-folded packing needs Cin in (2, 4, 8) or a multiple of 16", j0 + k);
        total = (int64_t)4 * CO * s.kd * s.kh * kxp;
      } else if (s.kind == PACK_XLINE) {
        const int CO = s.flip ? s.cin : s.cout, CI = s.flip ? s.cout : s.cin;
        B200_CHECK_ARG(((CO == 16 && (CI == 16 || CI == 48)) || (CO == 48 && CI == 16)) && s.kd == 3 && s.kh == 3 && s.kw == 3,
                       "pack_batch: job %d: x-line packing takes 3x3x3 kernels with (Cout', Cin') = (16, 16 | 48) or (48, 16)", j0 + k);
        total = (int64_t)27 * (CI / 16) * CO * 48;
      }
      d.total = total;
      // ~1 K elements per block, at most 512 blocks per job: the 1.8 M-element packs of the 256-channel layers must not become
      // the tail of the launch (64 blocks per job: 0.16 ms per launch, as slow as the 66 single launches it replaced)
      int64_t nb = ceil_div(total, 256 * 4);
      if (nb > 512) nb = 512;
      if (nb < 1) nb = 1;
      d.block_begin = block;
      d.n_blocks = (int)nb;
      block += (int)nb;
    }
    if (dtype == B200_BF16) pack_batch_table_kernel<__nv_bfloat16><<<block, 256, 0, st>>>(tab, n);
    else pack_batch_table_kernel<__half><<<block, 256, 0, st>>>(tab, n);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}

B200_EXPORT int b200_memset_zero(void* dst, int64_t bytes, void* stream) {
  B200_CHECK_ARG(dst && bytes >= 0, "memset_zero: bad args");
  if (bytes) B200_CUDA(cudaMemsetAsync(dst, 0, (size_t)bytes, (cudaStream_t)stream));
  return B200_OK;
}

B200_EXPORT int b200_write_floats(float* dst, const float* values, int32_t n, void* stream) {
  B200_CHECK_ARG(dst && values && n > 0 && n <= 16, "write_floats: 1..16 values");
  FloatBlock b;
  for (int i = 0; i < 16; ++i) b.v[i] = i < n ? values[i] : 0.f;
  write_floats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(dst, b, n);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_scale_by_dev(float* g, int64_t n, const double* denom, float mul, void* stream) {
  B200_CHECK_ARG(g && denom && n > 0, "scale_by_dev: bad args");
  scale_by_dev_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(g, n, denom, mul);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_sumsq(const float* g, int64_t n, double* out, void* stream) {
  B200_CHECK_ARG(g && out && n > 0, "sumsq: bad args");
  sumsq_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(g, n, out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
