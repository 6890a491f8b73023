// Cheaper arithmetic for the normalisation + activation kernels of the 16-bit engine.  Measured starting point (ncu, round 1):
// the c16 @128^3 launches execute ~11 (backward reduce / apply) and ~15 (forward apply) warp instructions per element slot with
// 57-63 % issue utilisation at 60-73 % of the copy bandwidth, and SiLU needs two MUFU operations per element (ex2 + rcp).
//   1. sigmoid(z) = 0.5 * tanh(0.5 z) + 0.5: ONE MUFU operation (tanh.approx.f32, relative error 2^-11 -- 16-bit dtypes only).
//   2. The backward evaluates the activation derivative once: the reduce pass leaves g = dy * act'(z) in place of dy (nothing
//      else reads dy afterwards) and the apply pass is three FMAs per element on (x, g).
// Same thread layouts, grids and coefficient tables as the product kernels (ops.cu: scale_shift_act_rows_kernel,
// norm_act_bwd_reduce_kernel, norm_act_bwd_apply_rows_kernel); tests/test_simt_emulation.py holds the chain built from these
// variants to the same double-precision formulas as the product chain.
#pragma once

__device__ __forceinline__ float tanh_approx(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return tanhf(x);
#endif
}
__device__ __forceinline__ float sigmoid_tanh(float z) { return fmaf(0.5f, tanh_approx(0.5f * z), 0.5f); }
__device__ __forceinline__ float silu_fwd_fast(float z) { return z * sigmoid_tanh(z); }
__device__ __forceinline__ float silu_grad_fast(float z) {
  const float s = sigmoid_tanh(z);
  return s * fmaf(z, 1.f - s, 1.f);
}

// forward apply: y = silu(x * scale + shift)
template <typename T, int VEC>
__global__ void scale_shift_silu_rows_fast_kernel(View<const T> x, View<T> y, const float* __restrict__ scale,
                                                  const float* __restrict__ shift, int cvn, int rows) {
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
        for (int i = 0; i < VEC; ++i) out.v[i] = from_f<T>(silu_fwd_fast(fmaf(to_f<T>(px[u].v[i]), sc[i], sh[i])));
        *reinterpret_cast<Pack<T, VEC>*>(yb + v * y.ld) = out;
      }
    }
  }
}

// backward pass 1: sums (sum g, sum g * (x - mean)) per (n, c) AND g written over dy.  The sums take g before it is rounded to T
// (they become dgamma / dbeta: on the emulator, sums of the rounded values were 0.3 % off on a 210-voxel tensor); the apply pass
// then subtracts means that differ from those of the stored values by the mean rounding error, far below one T ulp.
template <typename T, int VEC, int U, int MINB, bool WRITE_G = true>
__global__ void __launch_bounds__(256, MINB) norm_silu_bwd_reduce_g_kernel(View<const T> x, View<T> dy_g, const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd, int groups,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     double* __restrict__ red, int cv_count, int rows) {
  extern __shared__ double s_red[];
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int cv = tid % cv_count;
  const int row = tid / cv_count;
  const int cpg = x.c / groups;
  if (row < rows) {
    float ka[VEC], kb[VEC], mu_c[VEC], s[VEC], s2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = cv * VEC + i;
      const int g = c / cpg;
      const float mu = mean[(int64_t)n * groups + g], r = rstd[(int64_t)n * groups + g];
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
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
    T* db = dy_g.p + (int64_t)n * x.spatial * dy_g.ld + (int64_t)cv * VEC;
    for (int64_t v = v0 + row; v < v1; v += (int64_t)U * rows) {
      Pack<T, VEC> px[U], pd[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v + (int64_t)u * rows < v1) {
          px[u] = *reinterpret_cast<const Pack<T, VEC>*>(xb + (v + (int64_t)u * rows) * x.ld);
          pd[u] = *reinterpret_cast<const Pack<T, VEC>*>(db + (v + (int64_t)u * rows) * dy_g.ld);
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v + (int64_t)u * rows < v1) {
          Pack<T, VEC> pg;
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float fx = to_f<T>(px[u].v[i]);
            const float g = to_f<T>(pd[u].v[i]) * silu_grad_fast(fmaf(fx, ka[i], kb[i]));
            pg.v[i] = from_f<T>(g);
            s[i] += g;
            s2[i] = fmaf(g, fx - mu_c[i], s2[i]);
          }
          if (WRITE_G) *reinterpret_cast<Pack<T, VEC>*>(db + (v + (int64_t)u * rows) * dy_g.ld) = pg;
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

// backward pass 2 on (x, g): dx = g * k0 - x * P - Q, coefficients from norm_bwd_finalize_kernel unchanged
template <typename T, int VEC, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) norm_bwd_apply_g_rows_kernel(View<const T> x, View<const T> g, View<T> dx,
                                                                    const float* __restrict__ coef, int accumulate, int cvn, int rows) {
  const int n = blockIdx.y;
  const int cv = threadIdx.x % cvn, row = threadIdx.x / cvn;
  if (row >= rows) return;
  float k0[VEC], kp[VEC], kq[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float* cf = coef + ((int64_t)n * x.c + cv * VEC + i) * 4;
    k0[i] = cf[0]; kp[i] = cf[2]; kq[i] = cf[3];
  }
  const T* xb = x.p + (int64_t)n * x.spatial * x.ld + cv * VEC;
  const T* gb = g.p + (int64_t)n * x.spatial * g.ld + cv * VEC;
  T* ob = dx.p + (int64_t)n * x.spatial * dx.ld + cv * VEC;
  const int64_t stride = (int64_t)gridDim.x * rows;
  for (int64_t v0 = (int64_t)blockIdx.x * rows + row; v0 < x.spatial; v0 += U * stride) {
    Pack<T, VEC> px[U], pg[U], po[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < x.spatial) {
        px[u] = *reinterpret_cast<const Pack<T, VEC>*>(xb + v * x.ld);
        pg[u] = *reinterpret_cast<const Pack<T, VEC>*>(gb + v * g.ld);
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
          float r = fmaf(to_f<T>(pg[u].v[i]), k0[i], -fmaf(to_f<T>(px[u].v[i]), kp[i], kq[i]));
          if (accumulate) r += to_f<T>(po[u].v[i]);
          out.v[i] = from_f<T>(r);
        }
        *reinterpret_cast<Pack<T, VEC>*>(ob + v * dx.ld) = out;
      }
    }
  }
}


// backward pass 2 without the g hand-over: recomputes the derivative from (x, dy) with the one-MUFU sigmoid.  Pairs with
// norm_silu_bwd_reduce_g_kernel<..., WRITE_G = false>: no extra write in pass 1, one more tanh per element here.
template <typename T, int VEC, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) norm_silu_bwd_apply_fast_rows_kernel(View<const T> x, View<const T> dy, View<T> dx,
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
          const float g = to_f<T>(pd[u].v[i]) * silu_grad_fast(fmaf(fx, k0[i], kb[i]));
          float r = fmaf(g, k0[i], -fmaf(fx, kp[i], kq[i]));
          if (accumulate) r += to_f<T>(po[u].v[i]);
          out.v[i] = from_f<T>(r);
        }
        *reinterpret_cast<Pack<T, VEC>*>(ob + v * dx.ld) = out;
      }
    }
  }
}
