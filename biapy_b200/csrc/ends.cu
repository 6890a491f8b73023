// The two ends of the inference path: image normalisation in front of the first convolution
// (biapy/data/norm.py:44-220, 408-640), its inverse for the denoising workflow (norm.py:641-780) and the binarisation behind
// the merge (biapy/engine/semantic_seg.py:418-425, 524-531).  All HBM-bound streams: one read, one write.
#include "common.cuh"

namespace b200 {

constexpr int kMaxImgC = 16;

template <typename S> __device__ __forceinline__ float load_as_float(const S* p, int64_t i) { return (float)p[i]; }

// ----------------------------------------------------------------------------------------------- statistics
// per channel: min, max, sum, sum of squares (fp64 accumulation across threads, fp32 inside a thread's short run), is_binary
struct ClipParams { float on[kMaxImgC], lo[kMaxImgC], hi[kMaxImgC]; };

template <typename S>
__global__ void __launch_bounds__(256) image_stats_kernel(const S* __restrict__ src, int64_t voxels, int c, double* __restrict__ out,
                                                          const ClipParams cp) {
  // grid.y = channel; block-stride over voxels
  const int ch = blockIdx.y;
  float mn = INFINITY, mx = -INFINITY;
  double s1 = 0.0, s2 = 0.0;
  int bin = 1;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < voxels; v += (int64_t)gridDim.x * blockDim.x) {
    float x = load_as_float(src, v * c + ch);
    if (cp.on[ch] != 0.f) x = fminf(fmaxf(x, cp.lo[ch]), cp.hi[ch]);
    mn = fminf(mn, x); mx = fmaxf(mx, x);
    s1 += (double)x; s2 += (double)x * (double)x;
    bin &= (x == 0.f || x == 1.f) ? 1 : 0;
  }
  __shared__ float s_mn[8], s_mx[8];
  __shared__ double s_s1[8], s_s2[8];
  __shared__ int s_bin[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    bin &= __shfl_xor_sync(0xffffffffu, bin, o);
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; s_s1[warp] = s1; s_s2[warp] = s2; s_bin[warp] = bin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); s1 += s_s1[w]; s2 += s_s2[w]; bin &= s_bin[w]; }
    double* o = out + ch * 5;
    // min / max / is_binary through ordered-integer atomics on the double's bit pattern would need care with signs: use CAS loops
    unsigned long long* p;
    p = (unsigned long long*)(o + 0);
    for (unsigned long long old = *p;;) {
      if (__longlong_as_double((long long)old) <= (double)mn) break;
      const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong((double)mn));
      if (prev == old) break;
      old = prev;
    }
    p = (unsigned long long*)(o + 1);
    for (unsigned long long old = *p;;) {
      if (__longlong_as_double((long long)old) >= (double)mx) break;
      const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong((double)mx));
      if (prev == old) break;
      old = prev;
    }
    atomicAdd(o + 2, s1);
    atomicAdd(o + 3, s2);
    if (!bin) o[4] = 0.0;                      // racing writers all store the same value
  }
}

__global__ void image_stats_init_kernel(double* out, int c) {
  const int ch = threadIdx.x;
  if (ch < c) {
    out[ch * 5 + 0] = INFINITY; out[ch * 5 + 1] = -INFINITY; out[ch * 5 + 2] = 0.0; out[ch * 5 + 3] = 0.0; out[ch * 5 + 4] = 1.0;
  }
}

// ------------------------------------------------------------------------------------------------ normalise
struct NormParams { float clip[kMaxImgC], lo[kMaxImgC], hi[kMaxImgC], kind[kMaxImgC], a[kMaxImgC], b[kMaxImgC]; int c; };

template <typename S>
__global__ void __launch_bounds__(256) image_norm_kernel(const S* __restrict__ src, float* __restrict__ dst, int64_t total, const NormParams p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % p.c);
    float x = load_as_float(src, i);
    if (p.clip[ch] != 0.f) x = fminf(fmaxf(x, p.lo[ch]), p.hi[ch]);
    if (p.kind[ch] != 0.f) x = __fdiv_rn(__fsub_rn(x, p.a[ch]), p.b[ch]);
    dst[i] = x;
  }
}

struct DenormParams { int kind[kMaxImgC]; double a[kMaxImgC], b[kMaxImgC]; int c; };

template <typename D> __device__ __forceinline__ D store_from_double(double v, bool integer_range);
template <> __device__ __forceinline__ float store_from_double<float>(double v, bool) { return (float)v; }
template <> __device__ __forceinline__ uint8_t store_from_double<uint8_t>(double v, bool) { return (uint8_t)(long long)v; }
template <> __device__ __forceinline__ uint16_t store_from_double<uint16_t>(double v, bool) { return (uint16_t)(long long)v; }

template <typename D, int IMAX>
__global__ void __launch_bounds__(256) image_denorm_kernel(const float* __restrict__ src, D* __restrict__ dst, int64_t total, const DenormParams p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % p.c);
    float x = src[i];
    double v;
    if (p.kind[ch] == 1) {
      x = fminf(fmaxf(x, 0.f), 1.f);                          // np.clip(data, 0, 1) on the float32 array
      v = __dadd_rn(__dmul_rn((double)x, p.a[ch]), p.b[ch]);
    } else {
      v = __dadd_rn(__dmul_rn((double)x, p.a[ch]), p.b[ch]);
      if (IMAX > 0) v = fmin(fmax(rint(v), 0.0), (double)IMAX);   // np.round (half to even) then clip to the integer range
    }
    dst[i] = store_from_double<D>(v, IMAX > 0);
  }
}

__global__ void __launch_bounds__(256) binarize_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t n, float th) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i] > th ? 1 : 0;
}

template <typename D>
__global__ void __launch_bounds__(256) argmax_kernel(const float* __restrict__ src, D* __restrict__ dst, int64_t voxels, int c) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < voxels; v += (int64_t)gridDim.x * blockDim.x) {
    const float* row = src + v * c;
    float best = row[0];
    int arg = 0;
    bool nan = best != best;
    for (int k = 1; k < c && !nan; ++k) {
      const float x = row[k];
      if (x != x) { arg = k; nan = true; }                  // np.argmax returns the first NaN
      else if (x > best) { best = x; arg = k; }
    }
    dst[v] = (D)arg;
  }
}

// ------------------------------------------------------------------------------------------- order statistics
// np.percentile / kthvalue need the k-th smallest value of a channel exactly.  Radix select: the values map to 32-bit keys
// that sort like the values (unsigned integers as they are; floats with the sign bit flipped, negative floats inverted), and
// each pass histograms one digit of the keys that share the digits found so far.  One pass = one read of the channel.
template <typename S> __device__ __forceinline__ uint32_t select_key(S v) { return (uint32_t)v; }
template <> __device__ __forceinline__ uint32_t select_key<float>(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <typename S>
__global__ void __launch_bounds__(256) select_hist_kernel(const S* __restrict__ x, int64_t nvox, int c, int ch, int shift, int bits,
                                                          uint32_t prefix, int has_prefix, uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_h[2048];
  const int nb = 1 << bits;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) s_h[b] = 0u;
  __syncthreads();
  const uint32_t mask = (uint32_t)nb - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = select_key<S>(x[i * c + ch]);
    if (!has_prefix || (key >> (shift + bits)) == prefix) atomicAdd(&s_h[(key >> shift) & mask], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nb; b += blockDim.x)
    if (s_h[b]) atomicAdd(&hist[b], s_h[b]);
}

// np.histogram(x, bins=nbins) of a float32 array over its own [min, max] -- the equal-bin fast path of numpy
// (numpy/lib/_histograms_impl.py, "Fast algorithm for equal bins"): tentative index trunc(((x - first) / (last - first)) * nbins) in
// float32, nbins -> nbins - 1, then one step down if x < edges[i] and one step up if x >= edges[i + 1] (not from the last bin).
// `edges` (nbins + 1 float32) come from the host's own np.linspace, so the bins are the ones the caller's numpy would use.
// Every warp counts into a private copy of the histogram in shared memory.
__global__ void __launch_bounds__(256) edge_hist_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ edges, int nbins,
                                                        unsigned long long* __restrict__ counts) {
  extern __shared__ uint32_t s_eh[];               // [8 warps][nbins] counters, then nbins + 1 edges
  float* s_e = reinterpret_cast<float*>(s_eh + 8 * nbins);
  for (int b = threadIdx.x; b < 8 * nbins; b += blockDim.x) s_eh[b] = 0u;
  for (int b = threadIdx.x; b <= nbins; b += blockDim.x) s_e[b] = edges[b];
  __syncthreads();
  const float first = s_e[0], last = s_e[nbins];
  const float denom = __fsub_rn(last, first);
  uint32_t* mine = s_eh + (threadIdx.x >> 5) * nbins;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (!(v >= first && v <= last)) continue;      // numpy keeps first <= v <= last only (drops NaN)
    int idx = (int)__fmul_rn(__fdiv_rn(__fsub_rn(v, first), denom), (float)nbins);
    if (idx == nbins) idx -= 1;
    if (v < s_e[idx]) idx -= 1;
    if (v >= s_e[idx + 1] && idx != nbins - 1) idx += 1;
    atomicAdd(&mine[idx], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
    uint32_t t = 0;
    for (int w = 0; w < 8; ++w) t += s_eh[w * nbins + b];
    if (t) atomicAdd(&counts[b], (unsigned long long)t);
  }
}

static int stream_grid(int64_t total) {
  int64_t b = ceil_div(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace b200

#define B200_DISPATCH_IMG(dt, S, ...)                                              \
  switch (dt) {                                                                    \
    case B200_F32: { using S = float; __VA_ARGS__; break; }                        \
    case B200_U8: { using S = uint8_t; __VA_ARGS__; break; }                       \
    case B200_U16: { using S = uint16_t; __VA_ARGS__; break; }                     \
    default: b200::set_error("image dtype must be u8, u16 or f32, got %d", (int)(dt)); return B200_ERR_ARG; \
  }

B200_EXPORT int b200_image_stats(const void* src, int32_t dtype, int64_t voxels, int32_t c, const float* clip, double* out, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && out && voxels > 0 && c > 0 && c <= kMaxImgC, "image_stats: bad arguments (1 <= C <= %d)", kMaxImgC);
  ClipParams cp{};
  if (clip)
    for (int k = 0; k < c; ++k) { cp.on[k] = clip[3 * k]; cp.lo[k] = clip[3 * k + 1]; cp.hi[k] = clip[3 * k + 2]; }
  cudaStream_t st = (cudaStream_t)stream;
  image_stats_init_kernel<<<1, 32, 0, st>>>(out, c);
  int64_t bx = ceil_div(voxels, 256 * 8);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)c);
  B200_DISPATCH_IMG(dtype, S, { image_stats_kernel<S><<<grid, 256, 0, st>>>((const S*)src, voxels, c, out, cp); });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_image_norm_apply(const void* src, int32_t dtype, int64_t voxels, int32_t c, const float* params, float* dst,
                                      void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && dst && params && voxels > 0 && c > 0 && c <= kMaxImgC, "image_norm_apply: bad arguments (1 <= C <= %d)", kMaxImgC);
  NormParams p{};
  p.c = c;
  for (int k = 0; k < c; ++k) {
    const float* q = params + 6 * k;
    p.clip[k] = q[0]; p.lo[k] = q[1]; p.hi[k] = q[2]; p.kind[k] = q[3]; p.a[k] = q[4]; p.b[k] = q[5];
    B200_CHECK_ARG(q[3] == 0.f || q[3] == 1.f || q[3] == 2.f, "image_norm_apply: kind must be 0, 1 or 2");
  }
  const int64_t total = voxels * c;
  B200_DISPATCH_IMG(dtype, S, { image_norm_kernel<S><<<stream_grid(total), 256, 0, (cudaStream_t)stream>>>((const S*)src, dst, total, p); });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_image_denorm_apply(const float* src, int64_t voxels, int32_t c, const double* params, void* dst, int32_t dst_dtype,
                                        void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && dst && params && voxels > 0 && c > 0 && c <= kMaxImgC, "image_denorm_apply: bad arguments (1 <= C <= %d)", kMaxImgC);
  DenormParams p{};
  p.c = c;
  for (int k = 0; k < c; ++k) {
    p.kind[k] = (int)params[3 * k]; p.a[k] = params[3 * k + 1]; p.b[k] = params[3 * k + 2];
    B200_CHECK_ARG(p.kind[k] == 1 || p.kind[k] == 2, "image_denorm_apply: kind must be 1 (range) or 2 (mean / std)");
  }
  const int64_t total = voxels * c;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dst_dtype) {
    case B200_F32: image_denorm_kernel<float, 0><<<stream_grid(total), 256, 0, st>>>(src, (float*)dst, total, p); break;
    case B200_U8: image_denorm_kernel<uint8_t, 255><<<stream_grid(total), 256, 0, st>>>(src, (uint8_t*)dst, total, p); break;
    case B200_U16: image_denorm_kernel<uint16_t, 65535><<<stream_grid(total), 256, 0, st>>>(src, (uint16_t*)dst, total, p); break;
    default: B200_CHECK_ARG(false, "image_denorm_apply: dst dtype must be u8, u16 or f32");
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_binarize(const float* src, int64_t n, float threshold, uint8_t* dst, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && dst && n > 0, "binarize: bad arguments");
  binarize_kernel<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(src, dst, n, threshold);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_argmax_channels(const float* src, int64_t voxels, int32_t c, void* dst, int32_t dst_dtype, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && dst && voxels > 0 && c > 0, "argmax_channels: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dst_dtype == B200_U8) argmax_kernel<uint8_t><<<stream_grid(voxels), 256, 0, st>>>(src, (uint8_t*)dst, voxels, c);
  else if (dst_dtype == B200_U16) argmax_kernel<uint16_t><<<stream_grid(voxels), 256, 0, st>>>(src, (uint16_t*)dst, voxels, c);
  else B200_CHECK_ARG(false, "argmax_channels: dst dtype must be u8 or u16");
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_edge_hist(const float* src, int64_t n, const float* edges, int32_t nbins, uint64_t* counts, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && edges && counts && n > 0, "edge_hist: null pointer / empty input");
  B200_CHECK_ARG(nbins >= 1 && nbins <= 1024, "edge_hist: 1..1024 bins, got %d", nbins);
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * nbins, st));
  int64_t bx = ceil_div(n, 256 * 16);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  // a block never sees more than 2^32 elements of one bin: n / blocks < 2^32 for any tensor that fits the device
  const size_t smem = sizeof(uint32_t) * 8 * nbins + sizeof(float) * (nbins + 1);
  edge_hist_kernel<<<(unsigned)bx, 256, smem, st>>>(src, n, edges, nbins, (unsigned long long*)counts);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_select_hist(const void* src, int32_t dtype, int64_t voxels, int32_t c, int32_t ch, int32_t shift, int32_t bits,
                                 uint32_t prefix, int32_t has_prefix, uint32_t* hist, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && hist && voxels > 0 && c > 0 && ch >= 0 && ch < c, "select_hist: bad arguments");
  B200_CHECK_ARG(bits >= 1 && bits <= 11 && shift >= 0 && shift + bits <= 32 && (has_prefix == 0 || shift + bits < 32),
                 "select_hist: digit (shift %d, bits %d) out of range", shift, bits);
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) << bits, st));
  int64_t bx = ceil_div(voxels, 256 * 8);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  B200_DISPATCH_IMG(dtype, S, {
    select_hist_kernel<S><<<(unsigned)bx, 256, 0, st>>>((const S*)src, voxels, c, ch, shift, bits, prefix, has_prefix, hist);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}
