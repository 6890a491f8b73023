// Sliding-window bookkeeping and stitching: bit-exact patch-grid planner (host), crop gather and
// spline-weighted overlap-add gather (device).  Mirrors biapy/data/data_3D_manipulation.py:353-859 and the 2D
// twins in biapy/data/data_2D_manipulation.py:54-533 (see include/biapy_b200.h for the per-function map).
#include "common.cuh"


namespace b200 {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace b200

B200_EXPORT const char* b200_last_error(void) { return b200::g_err; }
B200_EXPORT int b200_version(void) { return 100; }

B200_EXPORT int b200_device_info(int device, int* sm_count, int* cc, int64_t* total_mem) {
  cudaDeviceProp p;
  B200_CUDA(cudaGetDeviceProperties(&p, device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc) *cc = p.major * 10 + p.minor;
  if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------ planner
static inline int64_t floordiv(int64_t a, int64_t b) {
  int64_t q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

B200_EXPORT int b200_plan_axis(int64_t dim, int64_t patch, int64_t pad, double overlap, b200_axis_plan* out) {
  B200_CHECK_ARG(out != nullptr, "plan_axis: null output");
  B200_CHECK_ARG(dim > 0 && patch > 0 && pad >= 0, "plan_axis: bad sizes dim=%lld patch=%lld pad=%lld",
                 (long long)dim, (long long)patch, (long long)pad);
  B200_CHECK_ARG(!(overlap >= 1.0 || overlap < 0.0), "'overlap' values must be floats between range [0, 1)");
  const int64_t core = patch - 2 * pad;
  // data_3D_manipulation.py:537-542 -- Python: int((P - 2p) * (1 - ov)); IEEE double product, truncation.
  // volatile keeps the compiler from contracting / re-associating the two double operations.
  volatile double keep = (overlap == 0.0) ? 1.0 : 1.0 - overlap;
  volatile double prod = (double)core * keep;
  int64_t step = (int64_t)prod;
  B200_CHECK_ARG(step != 0, "division by zero (step == 0 for patch=%lld pad=%lld overlap=%g)", (long long)patch,
                 (long long)pad, overlap);
  volatile double q = (double)dim / (double)step;                 // :543 math.ceil(dim / step)
  int64_t n = (int64_t)ceil(q);
  int64_t last = (n == 1) ? 0 : ((n - 1) * step + patch) - (dim + 2 * pad);   // :544
  int64_t per_block = (n > 1) ? floordiv(last, n - 1) : 0;        // :545
  step -= per_block;                                              // :546
  last -= per_block * (n - 1);                                    // :547
  out->dim = dim; out->patch = patch; out->pad = pad;
  out->step = step; out->n = n; out->last = last;
  out->core = core; out->ov_px = core - step;                     // :814-816
  return B200_OK;
}

B200_EXPORT int64_t b200_axis_start(const b200_axis_plan* p, int64_t i, int frame) {
  if (frame == 0) {  // crop: data_3D_manipulation.py:596-606 (padded frame, full patch extent)
    int64_t d = ((i * p->step + p->patch) < (p->dim + 2 * p->pad)) ? 0 : p->last;
    return i * p->step - d;
  }
  // merge: data_3D_manipulation.py:826-835 (original frame, core extent)
  int64_t d = ((i * p->step + p->core) < p->dim) ? 0 : p->last;
  return i * p->step - d;
}

B200_EXPORT int b200_spline_window_1d(int64_t size, int64_t ov_px, float* out) {
  B200_CHECK_ARG(size > 0 && out, "spline_window_1d: bad args");
  for (int64_t i = 0; i < size; ++i) out[i] = 1.0f;
  if (ov_px > 0) {
    int64_t ov = ov_px < size / 2 ? ov_px : size / 2;   // min(ov_pixels, size // 2)
    if (ov > 0) {
      // np.linspace(0, 1, ov + 2)[1:-1]: y_i = i * (1.0 / (ov + 1))     (numpy: arange * step + start)
      volatile double step = 1.0 / (double)(ov + 1);
      for (int64_t j = 0; j < ov; ++j) {
        volatile double x = (double)(j + 1) * step;
        volatile double x2 = x * x;                       // x ** 2
        volatile double omx = 1.0 - x;
        volatile double o2 = omx * omx;
        volatile double den = x2 + o2;
        den = den + 1e-8;
        volatile double t = x2 / den;
        float tf = (float)t;
        out[j] = tf;                                      // wind[:ov] = taper
        out[size - 1 - j] = tf;                           // wind[-ov:] = taper[::-1]
      }
    }
  }
  return B200_OK;
}

// ------------------------------------------------------------------------------------------- crop gather
namespace b200 {

__device__ __forceinline__ int64_t pad_src(int64_t j, int64_t dim, int mode) {
  // j is the coordinate in the un-padded frame (may be out of range); returns -1 for "constant zero"
  if (j >= 0 && j < dim) return j;
  switch (mode) {
    case B200_PAD_ZEROS: return -1;
    case B200_PAD_EDGE: return j < 0 ? 0 : dim - 1;
    case B200_PAD_REFLECT: {
      if (dim == 1) return 0;
      int64_t period = 2 * (dim - 1);
      int64_t m = j % period;
      if (m < 0) m += period;
      return m < dim ? m : period - m;
    }
    case B200_PAD_SYMMETRIC: {
      int64_t period = 2 * dim;
      int64_t m = j % period;
      if (m < 0) m += period;
      return m < dim ? m : period - 1 - m;
    }
    default: {  // wrap
      int64_t m = j % dim;
      if (m < 0) m += dim;
      return m;
    }
  }
}

struct CropParams {
  int64_t D, H, W, C, pd, ph, pw, nz, ny, nx, pad_z, pad_y, pad_x, total;
  int mode;
  int64_t src_z0, src_nz;      // src holds the planes [src_z0, src_z0 + src_nz) of the D-plane volume
};

constexpr int kMaxAxisPatches = 1024;

// grid = (chunks of ph*pw*C, pd, n_patches): only 32-bit index arithmetic per element
template <typename U>
__global__ void crop_gather_kernel(const U* __restrict__ src, U* __restrict__ dst, CropParams p,
                                   const int64_t* __restrict__ sz, const int64_t* __restrict__ sy,
                                   const int64_t* __restrict__ sx, int first) {
  const int patch = first + blockIdx.z;               // grid index of the patch; dst holds patches [first, first + gridDim.z)
  const int lz = blockIdx.y;
  const int ix = patch % (int)p.nx;
  const int iy = (patch / (int)p.nx) % (int)p.ny;
  const int iz = patch / (int)(p.nx * p.ny);
  int64_t z = pad_src(sz[iz] + lz - p.pad_z, p.D, p.mode);
  if (z >= 0) {
    z -= p.src_z0;
    if (z < 0 || z >= p.src_nz) z = -1;               // outside the shard the caller uploaded: never for planes_needed()
  }
  const int64_t y0 = sy[iy] - p.pad_y, x0 = sx[ix] - p.pad_x;
  const int plane = (int)(p.ph * p.pw * p.C);
  U* out = dst + ((int64_t)blockIdx.z * p.pd + lz) * plane;
  const int C = (int)p.C, pw = (int)p.pw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
    const int ch = i % C;
    const int t = i / C;
    const int lx = t % pw, ly = t / pw;
    const int64_t y = pad_src(y0 + ly, p.H, p.mode);
    const int64_t x = pad_src(x0 + lx, p.W, p.mode);
    U v = 0;
    if (z >= 0 && y >= 0 && x >= 0) v = src[((z * p.H + y) * p.W + x) * p.C + ch];
    out[i] = v;
  }
}

}  // namespace b200

static int crop_gather_impl(const void* src, int32_t dtype, int64_t D, int64_t H, int64_t W, int64_t C,
                            void* dst, int64_t pd, int64_t ph, int64_t pw,
                            const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                            const int64_t* starts_x, int64_t nx,
                            int64_t pad_z, int64_t pad_y, int64_t pad_x, int32_t pad_mode, int64_t first, int64_t count,
                            int64_t src_z0, int64_t src_nz, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && dst && starts_z && starts_y && starts_x, "crop_gather: null pointer");
  B200_CHECK_ARG(src_z0 >= 0 && src_nz > 0 && src_z0 + src_nz <= D, "crop_gather: source planes [%lld, %lld) outside the volume",
                 (long long)src_z0, (long long)(src_z0 + src_nz));
  B200_CHECK_ARG(valid_dtype(dtype) || dtype == 3, "crop_gather: bad dtype");
  B200_CHECK_ARG(nz > 0 && ny > 0 && nx > 0 && nz <= kMaxAxisPatches && ny <= kMaxAxisPatches && nx <= kMaxAxisPatches,
                 "crop_gather: patches per axis must be in [1, %d]", kMaxAxisPatches);
  B200_CHECK_ARG(pad_mode >= 0 && pad_mode <= 4, "crop_gather: bad pad mode");
  B200_CHECK_ARG(first >= 0 && count > 0 && first + count <= nz * ny * nx, "crop_gather: patch range [%lld, %lld) outside the grid of %lld",
                 (long long)first, (long long)(first + count), (long long)(nz * ny * nx));
  cudaStream_t st = (cudaStream_t)stream;
  CropParams p{D, H, W, C, pd, ph, pw, nz, ny, nx, pad_z, pad_y, pad_x, count * pd * ph * pw * C, pad_mode, src_z0, src_nz};
  int threads = 256;
  B200_CHECK_ARG(ph * pw * C < (1LL << 30) && nz * ny * nx <= 65535 * 32 && pd <= 65535, "crop_gather: patch too large");
  int64_t bx = ceil_div(ph * pw * C, threads);
  if (bx > 64) bx = 64;
  B200_CHECK_ARG(count <= 65535, "crop_gather: more than 65535 patches per call");
  dim3 blocks((unsigned)bx, (unsigned)pd, (unsigned)count);
  if (dtype == 3)
    crop_gather_kernel<uint8_t><<<blocks, threads, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, p,
                                                                     starts_z, starts_y, starts_x, (int)first);
  else if (dtype == B200_F32)
    crop_gather_kernel<uint32_t><<<blocks, threads, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, p,
                                                                      starts_z, starts_y, starts_x, (int)first);
  else
    crop_gather_kernel<uint16_t><<<blocks, threads, 0, st>>>((const uint16_t*)src, (uint16_t*)dst, p,
                                                                      starts_z, starts_y, starts_x, (int)first);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_crop_gather(const void* src, int32_t dtype, int64_t D, int64_t H, int64_t W, int64_t C,
                                 void* dst, int64_t pd, int64_t ph, int64_t pw,
                                 const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                                 const int64_t* starts_x, int64_t nx,
                                 int64_t pad_z, int64_t pad_y, int64_t pad_x, int32_t pad_mode, void* stream) {
  return crop_gather_impl(src, dtype, D, H, W, C, dst, pd, ph, pw, starts_z, nz, starts_y, ny, starts_x, nx, pad_z, pad_y, pad_x,
                          pad_mode, 0, nz * ny * nx, 0, D, stream);
}

B200_EXPORT int b200_crop_gather_range(const void* src, int32_t dtype, int64_t D, int64_t H, int64_t W, int64_t C,
                                       void* dst, int64_t pd, int64_t ph, int64_t pw,
                                       const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                                       const int64_t* starts_x, int64_t nx,
                                       int64_t pad_z, int64_t pad_y, int64_t pad_x, int32_t pad_mode, int64_t first, int64_t count,
                                       int64_t src_z0, int64_t src_nz, void* stream) {
  return crop_gather_impl(src, dtype, D, H, W, C, dst, pd, ph, pw, starts_z, nz, starts_y, ny, starts_x, nx, pad_z, pad_y, pad_x,
                          pad_mode, first, count, src_z0, src_nz, stream);
}

// ------------------------------------------------------------------------------------------ overlap-add
namespace b200 {

struct MergeParams {
  int64_t D, H, W, C, pz, py, px, pad_z, pad_y, pad_x, cz, cy, cx, nz, ny, nx, total;
};

// One thread per output element.  For every covering patch, in increasing patch index (z-major, then y,
// then x -- the order of the reference's triple loop, data_3D_manipulation.py:822-845):
//     acc  = acc  + (float(patch) * w)        w = (wz * wy) * wx   (float32, numpy left-to-right)
//     wsum = wsum + w
// and finally out = acc / (wsum + 1e-18f) cast to the output dtype (:849).  The explicit _rn intrinsics
// forbid FMA contraction, so float32 results are bit-identical to numpy's.
template <typename TI, typename TO>
__global__ void overlap_add_kernel(const TI* __restrict__ patches, TO* __restrict__ out, MergeParams p, int z0, int nz_out,
                                   const int64_t* __restrict__ sz, const int64_t* __restrict__ sy,
                                   const int64_t* __restrict__ sx, const float* __restrict__ wz,
                                   const float* __restrict__ wy, const float* __restrict__ wx) {
  extern __shared__ int64_t s_starts[];  // nz + ny + nx
  int64_t* s_z = s_starts;
  int64_t* s_y = s_z + p.nz;
  int64_t* s_x = s_y + p.ny;
  for (int i = threadIdx.x; i < p.nz; i += blockDim.x) s_z[i] = sz[i];
  for (int i = threadIdx.x; i < p.ny; i += blockDim.x) s_y[i] = sy[i];
  for (int i = threadIdx.x; i < p.nx; i += blockDim.x) s_x[i] = sx[i];
  __syncthreads();
  const int plane = (int)(p.H * p.W * p.C);
  const int C = (int)p.C, Wd = (int)p.W;
  for (int64_t zl = blockIdx.y; zl < nz_out; zl += gridDim.y)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
    const int64_t z = z0 + zl;
    const int ch = i % C;
    const int tq = i / C;
    const int64_t x = tq % Wd, y = tq / Wd;
    const int64_t idx = zl * plane + i;
    float acc = 0.f, wsum = 0.f;
    for (int64_t iz = 0; iz < p.nz; ++iz) {
      int64_t lz = z - s_z[iz];
      if (lz < 0 || lz >= p.cz) continue;
      float fz = wz[lz];
      for (int64_t iy = 0; iy < p.ny; ++iy) {
        int64_t ly = y - s_y[iy];
        if (ly < 0 || ly >= p.cy) continue;
        float fzy = __fmul_rn(fz, wy[ly]);
        for (int64_t ix = 0; ix < p.nx; ++ix) {
          int64_t lx = x - s_x[ix];
          if (lx < 0 || lx >= p.cx) continue;
          float w = __fmul_rn(fzy, wx[lx]);
          int64_t c = (iz * p.ny + iy) * p.nx + ix;
          int64_t off = (((c * p.pz + lz + p.pad_z) * p.py + ly + p.pad_y) * p.px + lx + p.pad_x) * p.C + ch;
          float v = to_f<TI>(patches[off]);
          acc = __fadd_rn(acc, __fmul_rn(v, w));
          wsum = __fadd_rn(wsum, w);
        }
      }
    }
    out[idx] = from_f<TO>(__fdiv_rn(acc, __fadd_rn(wsum, 1e-18f)));
  }
}

// Cover-table variant (the default): every block first tabulates, for all x and all y of the plane, WHICH patches of that
// axis cover the coordinate -- a 64-bit mask per coordinate in shared memory (W*nx + H*ny compares spread over the block; the
// covering set is not always an index range: with padding and overlap the reference's merge starts are not monotone, e.g.
// [0, 3, 6, 9, 12, 10, 13]) -- and each output element then visits exactly its covering patches in increasing index order
// with 32-bit arithmetic, instead of walking the nz x ny x nx loop nest of the kernel above with 64-bit range checks at every
// level.  Same operation order, bit-identical results; used when no axis has more than 64 patches.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) overlap_add_cover_kernel(const TI* __restrict__ patches, TO* __restrict__ out, MergeParams p,
                                         int z0, int nz_out, const int64_t* __restrict__ sz, const int64_t* __restrict__ sy,
                                         const int64_t* __restrict__ sx, const float* __restrict__ wz,
                                         const float* __restrict__ wy, const float* __restrict__ wx) {
  extern __shared__ unsigned long long s_mask[];             // x masks [W], y masks [H], then the starts as int
  const int nz = (int)p.nz, ny = (int)p.ny, nx = (int)p.nx, H = (int)p.H, W = (int)p.W, C = (int)p.C;
  const int cz = (int)p.cz, cy = (int)p.cy, cx = (int)p.cx;
  unsigned long long* x_mask = s_mask;
  unsigned long long* y_mask = x_mask + W;
  int* s_z = reinterpret_cast<int*>(y_mask + H);
  int* s_y = s_z + nz;
  int* s_x = s_y + ny;
  for (int i = threadIdx.x; i < nz; i += blockDim.x) s_z[i] = (int)sz[i];
  for (int i = threadIdx.x; i < ny; i += blockDim.x) s_y[i] = (int)sy[i];
  for (int i = threadIdx.x; i < nx; i += blockDim.x) s_x[i] = (int)sx[i];
  __syncthreads();
  for (int c = threadIdx.x; c < W + H; c += blockDim.x) {
    const bool isx = c < W;
    const int coord = isx ? c : c - W, n = isx ? nx : ny, core = isx ? cx : cy;
    const int* st = isx ? s_x : s_y;
    unsigned long long m = 0;
    for (int i = 0; i < n; ++i) {
      const int l = coord - st[i];
      if (l >= 0 && l < core) m |= 1ull << i;
    }
    (isx ? x_mask : y_mask)[coord] = m;
  }
  __syncthreads();
  const int plane = H * W * C;
  const int64_t pvol = p.pz * p.py * p.px * p.C;           // elements per patch
  const int row_el = (int)(p.px * p.C);
  for (int zl = blockIdx.y; zl < nz_out; zl += gridDim.y) {
    const int z = z0 + zl;
    unsigned long long zm = 0;
    for (int i = 0; i < nz; ++i) {
      const int l = z - s_z[i];
      if (l >= 0 && l < cz) zm |= 1ull << i;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
      const int ch = i % C;
      const int tq = i / C;
      const int x = tq % W, y = tq / W;
      const unsigned long long xm = x_mask[x], ym = y_mask[y];
      float acc = 0.f, wsum = 0.f;
      for (unsigned long long a = zm; a; a &= a - 1) {
        const int iz = __ffsll((long long)a) - 1, lz = z - s_z[iz];
        const float fz = __ldg(wz + lz);
        for (unsigned long long b = ym; b; b &= b - 1) {
          const int iy = __ffsll((long long)b) - 1, ly = y - s_y[iy];
          const float fzy = __fmul_rn(fz, __ldg(wy + ly));
          // element offset of (patch (iz, iy, 0), row lz, ly) -- 64-bit once per (z, y) pair
          const TI* rowp = patches + ((int64_t)(iz * ny + iy) * nx) * pvol +
                           ((int64_t)(lz + (int)p.pad_z) * p.py + (ly + (int)p.pad_y)) * row_el + ch;
          for (unsigned long long c = xm; c; c &= c - 1) {
            const int ix = __ffsll((long long)c) - 1, lx = x - s_x[ix];
            const float w = __fmul_rn(fzy, __ldg(wx + lx));
            const float v = to_f<TI>(rowp[(int64_t)ix * pvol + (lx + (int)p.pad_x) * C]);
            acc = __fadd_rn(acc, __fmul_rn(v, w));
            wsum = __fadd_rn(wsum, w);
          }
        }
      }
      out[(int64_t)zl * plane + i] = from_f<TO>(__fdiv_rn(acc, __fadd_rn(wsum, 1e-18f)));
    }
  }
}

// Slot variant (default for the usual grids, overlap < 50 %): with at most TWO covering patches per axis the covering set of an
// output element is a 2 x 2 x 2 box of slots, so the walk over the covering patches becomes eight predicated, fully unrolled steps
// whose loads are independent of each other -- the cover kernel above chases mask bits one patch at a time, exposes one load
// latency per patch and spends ~280 instructions per element on bookkeeping.  Here the bookkeeping is hoisted: a block owns one
// output plane z (slots and weights of z are block constants), a band of y rows (their slots, weights and patch-row offsets sit
// in a shared table built once per block) and every thread keeps the slots of ITS x column in registers while it walks down the
// band.  The operation order (z-major, then y, then x: increasing patch index) and every float32 operation are those of
// overlap_add_kernel: bit-identical.  Elements with three or more covering patches on some axis take a plain range walk.
// `z0`: first output plane of the slab this launch produces (sharded inference merges one z slab per rank).
// 32-bit element offsets: the launcher checks n_patches * patch volume < 2^31.
struct SlotRow { int n, off0, off1; float f0, f1; };      // covering patches of one coordinate (n > 2: walk), row offsets, weights

// NZ x NY x (1 or 2) covering patches of one output element, in patch order (z-major, y, then x).  NZ / NY are compile-time (they
// are the same for a whole warp), the second x slot is a per-thread predicate.  All loads are issued before the first add.
template <typename TI, int NZ, int NY>
__device__ __forceinline__ void slot_accumulate(const TI* __restrict__ b0, const TI* __restrict__ b1, bool two_x, const int (&zoff)[2],
                                                const float (&fz)[2], const SlotRow& t, float fx0, float fx1, float& acc, float& wsum) {
  float v0[NZ * NY], v1[NZ * NY];
#pragma unroll
  for (int kz = 0; kz < NZ; ++kz)
#pragma unroll
    for (int ky = 0; ky < NY; ++ky) {
      const int o = zoff[kz] + (ky ? t.off1 : t.off0);
      v0[kz * NY + ky] = to_f<TI>(b0[o]);
      v1[kz * NY + ky] = two_x ? to_f<TI>(b1[o]) : 0.f;
    }
#pragma unroll
  for (int kz = 0; kz < NZ; ++kz)
#pragma unroll
    for (int ky = 0; ky < NY; ++ky) {
      const float fzy = __fmul_rn(fz[kz], ky ? t.f1 : t.f0);
      const float w0 = __fmul_rn(fzy, fx0);
      acc = __fadd_rn(acc, __fmul_rn(v0[kz * NY + ky], w0));
      wsum = __fadd_rn(wsum, w0);
      if (two_x) {
        const float w1 = __fmul_rn(fzy, fx1);
        acc = __fadd_rn(acc, __fmul_rn(v1[kz * NY + ky], w1));
        wsum = __fadd_rn(wsum, w1);
      }
    }
}

template <typename TI, typename TO, int UNROLL, int MINB>
__global__ void __launch_bounds__(256, MINB) overlap_add_slot_kernel(const TI* __restrict__ patches, TO* __restrict__ out, MergeParams p,
                                         int z0, int nz_out, int rows_per_block,
                                         const int64_t* __restrict__ sz, const int64_t* __restrict__ sy,
                                         const int64_t* __restrict__ sx, const float* __restrict__ wz,
                                         const float* __restrict__ wy, const float* __restrict__ wx) {
  extern __shared__ unsigned long long s_raw[];
  const int nz = (int)p.nz, ny = (int)p.ny, nx = (int)p.nx, H = (int)p.H, W = (int)p.W, C = (int)p.C;
  const int cz = (int)p.cz, cy = (int)p.cy, cx = (int)p.cx;
  SlotRow* s_row = reinterpret_cast<SlotRow*>(s_raw);                 // [rows_per_block]
  int* s_z = reinterpret_cast<int*>(s_row + rows_per_block);
  int* s_y = s_z + nz;
  int* s_x = s_y + ny;
  const int pvol = (int)(p.pz * p.py * p.px * p.C);          // elements per patch
  const int prow = (int)(p.px * p.C);                        // elements of one patch x row
  const int padz = (int)p.pad_z, pady = (int)p.pad_y, padx = (int)p.pad_x;
  const int row_el = W * C;
  for (int i = threadIdx.x; i < nz; i += blockDim.x) s_z[i] = (int)sz[i];
  for (int i = threadIdx.x; i < ny; i += blockDim.x) s_y[i] = (int)sy[i];
  for (int i = threadIdx.x; i < nx; i += blockDim.x) s_x[i] = (int)sx[i];
  __syncthreads();
  const int y_begin = blockIdx.x * rows_per_block;
  const int y_end = H < y_begin + rows_per_block ? H : y_begin + rows_per_block;
  for (int r = threadIdx.x; r < y_end - y_begin; r += blockDim.x) {
    const int y = y_begin + r;
    SlotRow t{0, 0, 0, 0.f, 0.f};
    for (int i = 0; i < ny; ++i) {
      const int l = y - s_y[i];
      if (l >= 0 && l < cy) {
        const int off = i * nx * pvol + (l + pady) * prow;
        if (t.n == 0) { t.off0 = off; t.f0 = wy[l]; }
        else if (t.n == 1) { t.off1 = off; t.f1 = wy[l]; }
        ++t.n;
      }
    }
    s_row[r] = t;
  }
  __syncthreads();
  for (int zl = blockIdx.y; zl < nz_out; zl += gridDim.y) {
    const int z = z0 + zl;
    int nzc = 0, zoff[2] = {0, 0};
    float fz[2] = {0.f, 0.f};
    for (int i = 0; i < nz; ++i) {
      const int l = z - s_z[i];
      if (l >= 0 && l < cz) {
        if (nzc < 2) { zoff[nzc] = i * ny * nx * pvol + (l + padz) * (int)p.py * prow; fz[nzc] = __ldg(wz + l); }
        ++nzc;
      }
    }
    for (int e0 = threadIdx.x; e0 < row_el; e0 += blockDim.x) {
      const int x = e0 / C, ch = e0 - x * C;
      int nxc = 0, xoff[2] = {0, 0};
      float fx[2] = {0.f, 0.f};
      for (int i = 0; i < nx; ++i) {
        const int l = x - s_x[i];
        if (l >= 0 && l < cx) {
          if (nxc < 2) { xoff[nxc] = i * pvol + (l + padx) * C + ch; fx[nxc] = __ldg(wx + l); }
          ++nxc;
        }
      }
      const bool zx_fast = (nzc == 1 || nzc == 2) && (nxc == 1 || nxc == 2);
      const bool two_x = nxc == 2;
      const TI* b0 = patches + xoff[0];
      const TI* b1 = patches + xoff[1];
      TO* op = out + ((int64_t)zl * H + y_begin) * row_el + e0;
#pragma unroll UNROLL
      for (int r = 0; r < y_end - y_begin; ++r) {
        const SlotRow t = s_row[r];
        float acc = 0.f, wsum = 0.f;
        if (zx_fast && (t.n == 1 || t.n == 2)) {
          if (nzc == 1) {
            if (t.n == 1) slot_accumulate<TI, 1, 1>(b0, b1, two_x, zoff, fz, t, fx[0], fx[1], acc, wsum);
            else slot_accumulate<TI, 1, 2>(b0, b1, two_x, zoff, fz, t, fx[0], fx[1], acc, wsum);
          } else {
            if (t.n == 1) slot_accumulate<TI, 2, 1>(b0, b1, two_x, zoff, fz, t, fx[0], fx[1], acc, wsum);
            else slot_accumulate<TI, 2, 2>(b0, b1, two_x, zoff, fz, t, fx[0], fx[1], acc, wsum);
          }
        } else {
          const int y = y_begin + r;
          for (int iz = 0; iz < nz; ++iz) {
            const int lz = z - s_z[iz];
            if (lz < 0 || lz >= cz) continue;
            const float gz = __ldg(wz + lz);
            for (int iy = 0; iy < ny; ++iy) {
              const int ly = y - s_y[iy];
              if (ly < 0 || ly >= cy) continue;
              const float gzy = __fmul_rn(gz, __ldg(wy + ly));
              for (int ix = 0; ix < nx; ++ix) {
                const int lx = x - s_x[ix];
                if (lx < 0 || lx >= cx) continue;
                const float w = __fmul_rn(gzy, __ldg(wx + lx));
                const int64_t c = ((int64_t)iz * ny + iy) * nx + ix;
                const float val = to_f<TI>(patches[c * pvol + ((int64_t)(lz + padz) * p.py + (ly + pady)) * prow + (lx + padx) * C + ch]);
                acc = __fadd_rn(acc, __fmul_rn(val, w));
                wsum = __fadd_rn(wsum, w);
              }
            }
          }
        }
        op[(int64_t)r * row_el] = from_f<TO>(__fdiv_rn(acc, __fadd_rn(wsum, 1e-18f)));
      }
    }
  }
}

template <typename TI, typename TO>
static int launch_overlap_add(const void* patches, void* out, const MergeParams& p, int64_t z0, int64_t nz_out, const int64_t* sz,
                              const int64_t* sy, const int64_t* sx, const float* wz, const float* wy,
                              const float* wx, cudaStream_t st) {
  int threads = 256;
  B200_CHECK_ARG(p.H * p.W * p.C < (1LL << 31), "overlap_add: plane too large");
  B200_CHECK_ARG(z0 >= 0 && nz_out > 0 && z0 + nz_out <= p.D, "overlap_add: slab [%lld, %lld) outside the volume (%lld planes)",
                 (long long)z0, (long long)(z0 + nz_out), (long long)p.D);
  int64_t bx = ceil_div(p.H * p.W * p.C, threads);
  if (bx > 1024) bx = 1024;
  const unsigned gy = (unsigned)(nz_out < 65535 ? nz_out : 65535);
  dim3 blocks((unsigned)bx, gy);
  // cover masks in shared memory (64 patches per axis at most), 32-bit coordinates, <= 48 KB
  const size_t tab = sizeof(unsigned long long) * (size_t)(p.W + p.H) + sizeof(int) * (size_t)(p.nz + p.ny + p.nx);
  // B200_MERGE_KERNEL: slot (default) | cover | plain
  static const int variant = [] {
    const char* e = getenv("B200_MERGE_KERNEL");
    if (getenv("B200_MERGE_COVER") && strcmp(getenv("B200_MERGE_COVER"), "0") == 0) return 2;
    return !e ? 0 : !strcmp(e, "cover") ? 1 : !strcmp(e, "plain") ? 2 : 0;
  }();
  if (variant != 2 && tab <= 48 * 1024 && p.nz <= 64 && p.ny <= 64 && p.nx <= 64 && p.D < (1LL << 30) && p.px * p.C < (1LL << 30)) {
    if (variant == 0 && (p.nz * p.ny * p.nx) * (p.pz * p.py * p.px * p.C) < (1LL << 31)) {
      static const int rows_env = getenv("B200_MERGE_ROWS") ? atoi(getenv("B200_MERGE_ROWS")) : 0;     // y rows per block
      static const int unroll2 = getenv("B200_MERGE_UNROLL") ? atoi(getenv("B200_MERGE_UNROLL")) : 2;
      // a block = one z plane x a band of y rows; about 8 blocks per SM over the whole launch, bands of 16..128 rows
      int64_t bands = ceil_div((int64_t)sm_count() * 8, (int64_t)gy);
      int64_t rpb = rows_env > 0 ? rows_env : ceil_div(p.H, bands < 1 ? 1 : bands);
      if (rows_env <= 0) { if (rpb < 16) rpb = 16; if (rpb > 128) rpb = 128; }
      if (rpb > p.H) rpb = p.H;
      const dim3 g((unsigned)ceil_div(p.H, rpb), gy);
      const size_t smem = sizeof(SlotRow) * (size_t)rpb + sizeof(int) * (size_t)(p.nz + p.ny + p.nx) + 16;
      static const int occ6 = getenv("B200_MERGE_OCC") ? atoi(getenv("B200_MERGE_OCC")) == 6 : 0;   // 6 blocks / SM (40 registers, spills)
#define B200_SLOT(U, MB) overlap_add_slot_kernel<TI, TO, U, MB><<<g, threads, smem, st>>>((const TI*)patches, (TO*)out, p, (int)z0, \
                                                                                      (int)nz_out, (int)rpb, sz, sy, sx, wz, wy, wx)
      if (occ6) B200_SLOT(2, 6); else if (unroll2 == 1) B200_SLOT(1, 5); else if (unroll2 == 4) B200_SLOT(4, 5); else B200_SLOT(2, 5);
#undef B200_SLOT
      B200_LAUNCH_CHECK();
      return B200_OK;
    }
    // few, fat blocks per plane: the table build is per block
    int64_t bxc = ceil_div(p.H * p.W * p.C, (int64_t)threads * 8);
    if (bxc > 64) bxc = 64;
    if (bxc < 1) bxc = 1;
    dim3 bc((unsigned)bxc, gy);
    overlap_add_cover_kernel<TI, TO><<<bc, threads, tab, st>>>((const TI*)patches, (TO*)out, p, (int)z0, (int)nz_out, sz, sy, sx, wz,
                                                              wy, wx);
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  size_t smem = sizeof(int64_t) * (p.nz + p.ny + p.nx);
  overlap_add_kernel<TI, TO><<<blocks, threads, smem, st>>>((const TI*)patches, (TO*)out, p, (int)z0, (int)nz_out, sz, sy, sx,
                                                                     wz, wy, wx);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace b200

static int overlap_add_impl(const void* patches, int32_t dtype_in, void* out, int32_t dtype_out,
                            int64_t D, int64_t H, int64_t W, int64_t C,
                            int64_t pz, int64_t py, int64_t px, int64_t pad_z, int64_t pad_y, int64_t pad_x,
                            const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                            const int64_t* starts_x, int64_t nx,
                            const float* win_z, const float* win_y, const float* win_x, int64_t z0, int64_t nz_out, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(patches && out && starts_z && starts_y && starts_x && win_z && win_y && win_x,
                 "overlap_add: null pointer");
  B200_CHECK_ARG(valid_dtype(dtype_in) && valid_dtype(dtype_out), "overlap_add: bad dtype");
  B200_CHECK_ARG(nz > 0 && ny > 0 && nx > 0 && nz + ny + nx <= 4096, "overlap_add: bad patch counts");
  B200_CHECK_ARG(pz > 2 * pad_z && py > 2 * pad_y && px > 2 * pad_x, "overlap_add: padding too large");
  MergeParams p{D, H, W, C, pz, py, px, pad_z, pad_y, pad_x, pz - 2 * pad_z, py - 2 * pad_y, px - 2 * pad_x,
                nz, ny, nx, D * H * W * C};
  cudaStream_t st = (cudaStream_t)stream;
#define OA(TI, TO) return launch_overlap_add<TI, TO>(patches, out, p, z0, nz_out, starts_z, starts_y, starts_x, win_z, win_y, win_x, st)
  if (dtype_in == B200_F32 && dtype_out == B200_F32) OA(float, float);
  if (dtype_in == B200_F16 && dtype_out == B200_F16) OA(__half, __half);
  if (dtype_in == B200_F16 && dtype_out == B200_F32) OA(__half, float);
  if (dtype_in == B200_BF16 && dtype_out == B200_BF16) OA(__nv_bfloat16, __nv_bfloat16);
  if (dtype_in == B200_BF16 && dtype_out == B200_F32) OA(__nv_bfloat16, float);
  if (dtype_in == B200_F32 && dtype_out == B200_F16) OA(float, __half);
  if (dtype_in == B200_F32 && dtype_out == B200_BF16) OA(float, __nv_bfloat16);
#undef OA
  set_error("overlap_add: unsupported dtype pair %d -> %d", dtype_in, dtype_out);
  return B200_ERR_UNSUPPORTED;
}

B200_EXPORT int b200_overlap_add(const void* patches, int32_t dtype_in, void* out, int32_t dtype_out,
                                 int64_t D, int64_t H, int64_t W, int64_t C,
                                 int64_t pz, int64_t py, int64_t px, int64_t pad_z, int64_t pad_y, int64_t pad_x,
                                 const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                                 const int64_t* starts_x, int64_t nx,
                                 const float* win_z, const float* win_y, const float* win_x, void* stream) {
  return overlap_add_impl(patches, dtype_in, out, dtype_out, D, H, W, C, pz, py, px, pad_z, pad_y, pad_x, starts_z, nz, starts_y, ny,
                          starts_x, nx, win_z, win_y, win_x, 0, D, stream);
}

B200_EXPORT int b200_overlap_add_slab(const void* patches, int32_t dtype_in, void* out, int32_t dtype_out,
                                      int64_t D, int64_t H, int64_t W, int64_t C,
                                      int64_t pz, int64_t py, int64_t px, int64_t pad_z, int64_t pad_y, int64_t pad_x,
                                      const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                                      const int64_t* starts_x, int64_t nx,
                                      const float* win_z, const float* win_y, const float* win_x, int64_t z0, int64_t nz_out,
                                      void* stream) {
  return overlap_add_impl(patches, dtype_in, out, dtype_out, D, H, W, C, pz, py, px, pad_z, pad_y, pad_x, starts_z, nz, starts_y, ny,
                          starts_x, nx, win_z, win_y, win_x, z0, nz_out, stream);
}

// ---------------------------------------------------------------------------------------- by-chunks tile grid
// chunked_test_pair_data_generator.py:276-295 (plain integer arithmetic: math.ceil(a / b) on Python ints is evaluated
// through a double division there; every value on this path is far below 2^53, so the integer ceil is identical).
static inline int64_t ceil_div_i(int64_t a, int64_t b) { return (a + b - 1) / b; }

B200_EXPORT int b200_chunk_grid_plan(const int64_t dim[3], const int64_t crop[3], const int64_t pad[3], int64_t z_start,
                                     int64_t z_end, b200_chunk_grid* g) {
  B200_CHECK_ARG(dim && crop && pad && g, "chunk_grid_plan: null pointer");
  static const char* ax = "ZYX";
  for (int a = 0; a < 3; ++a) {
    B200_CHECK_ARG(dim[a] > 0 && crop[a] > 0 && pad[a] >= 0, "chunk_grid_plan: bad sizes on axis %c", ax[a]);
    B200_CHECK_ARG(crop[a] <= dim[a], "%c Axis problem: %lld greater than %lld (you can reduce 'DATA.PATCH_SIZE' in that axis)",
                   ax[a], (long long)crop[a], (long long)dim[a]);
  }
  for (int a = 0; a < 3; ++a)
    B200_CHECK_ARG(pad[a] < crop[a] / 2, "'Padding' can not be greater than half of 'crop_shape'. Max value for the given input "
                   "shape is (%lld, %lld, %lld)", (long long)(crop[0] / 2 - 1), (long long)(crop[1] / 2 - 1), (long long)(crop[2] / 2 - 1));
  for (int a = 0; a < 3; ++a) {
    g->dim[a] = dim[a]; g->crop[a] = crop[a]; g->pad[a] = pad[a];
    g->step[a] = crop[a] - 2 * pad[a];
    g->vols[a] = ceil_div_i(dim[a], g->step[a]);
  }
  const int64_t ez0 = (z_start == -1) ? 0 : z_start, ez1 = (z_end == -1) ? dim[0] : z_end;
  B200_CHECK_ARG(ez0 >= 0 && ez1 >= 0, "chunk_grid_plan: negative Z range");
  g->z_vol_start = ceil_div_i(ez0, g->step[0]);
  const int64_t zend = ceil_div_i(ez1, g->step[0]);
  g->z_vol_end = zend < g->vols[0] ? zend : g->vols[0];
  g->total = (g->z_vol_end - g->z_vol_start) * g->vols[1] * g->vols[2];
  return B200_OK;
}

B200_EXPORT int b200_chunk_patch_coords(const b200_chunk_grid* g, int64_t vol_id, int64_t out[27]) {
  B200_CHECK_ARG(g && out, "chunk_patch_coords: null pointer");
  B200_CHECK_ARG(vol_id >= 0 && vol_id < g->total, "chunk_patch_coords: tile %lld outside the grid of %lld", (long long)vol_id,
                 (long long)g->total);
  int64_t pos[3];
  pos[2] = vol_id % g->vols[2];
  pos[1] = (vol_id / g->vols[2]) % g->vols[1];
  pos[0] = vol_id / (g->vols[2] * g->vols[1]) + g->z_vol_start;
  for (int a = 0; a < 3; ++a) {
    out[a] = pos[a];
    const int64_t lo = pos[a] * g->step[a] - g->pad[a], hi = (pos[a] + 1) * g->step[a] + g->pad[a];
    const int64_t es = lo > 0 ? lo : 0, ee = hi < g->dim[a] ? hi : g->dim[a];                      // :463-470
    out[3 + 2 * a] = es; out[4 + 2 * a] = ee;
    out[9 + 2 * a] = pos[a] * g->step[a];                                                          // :474-481
    out[10 + 2 * a] = (pos[a] + 1) * g->step[a] < g->dim[a] ? (pos[a] + 1) * g->step[a] : g->dim[a];
    const int64_t left = lo < 0 ? -lo : 0, right = g->crop[a] - (ee - es) - left;                    // :536-541
    out[15 + 2 * a] = left; out[16 + 2 * a] = right;
    out[21 + 2 * a] = left > g->pad[a] ? left : g->pad[a];                                          // :555-560
    out[22 + 2 * a] = right > g->pad[a] ? right : g->pad[a];
  }
  return B200_OK;
}

namespace b200 {

__device__ __forceinline__ int reflect_in(int j, int len) {      // numpy 'reflect' (edge not repeated), any reach
  if (len == 1) return 0;
  const int period = 2 * (len - 1);
  int m = j % period;
  if (m < 0) m += period;
  return m < len ? m : period - m;
}

// grid = (chunks of a (ph, pw, C) plane, pd, n tiles).  Rows of `pw * C` elements are contiguous on both sides wherever the
// x window is not reflected, so consecutive threads read / write consecutive addresses.
template <typename U>
__global__ void chunk_extract_kernel(const U* __restrict__ src, U* __restrict__ dst, const int64_t* __restrict__ desc, int H, int W,
                                     int C, int pd, int ph, int pw) {
  const int t = blockIdx.z, lz = blockIdx.y;
  const int64_t* d = desc + (int64_t)t * 9;
  const int z0 = (int)d[0], zl = (int)d[1], zp = (int)d[2], y0 = (int)d[3], yl = (int)d[4], yp = (int)d[5];
  const int x0 = (int)d[6], xl = (int)d[7], xp = (int)d[8];
  const int64_t z = z0 + reflect_in(lz - zp, zl);
  const int plane = ph * pw * C;
  U* out = dst + ((int64_t)t * pd + lz) * plane;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
    const int ch = i % C, q = i / C;
    const int lx = q % pw, ly = q / pw;
    const int64_t y = y0 + reflect_in(ly - yp, yl), x = x0 + reflect_in(lx - xp, xl);
    out[i] = src[((z * H + y) * W + x) * C + ch];
  }
}

template <typename TI, typename TO>
__global__ void chunk_insert_kernel(const TI* __restrict__ patches, TO* __restrict__ out, const int64_t* __restrict__ desc, int H, int W,
                                    int C, int pd, int ph, int pw, int mode) {
  const int t = blockIdx.z;
  const int64_t* d = desc + (int64_t)t * 9;
  const int oz = (int)d[0], sz = (int)d[1], cz = (int)d[2], oy = (int)d[3], sy = (int)d[4], cy = (int)d[5];
  const int ox = (int)d[6], sx = (int)d[7], cx = (int)d[8];
  const int row = sx * C;                       // contiguous in the patch (from cx) and in the volume (from ox)
  const int plane = sy * row;
  for (int lz = blockIdx.y; lz < sz; lz += gridDim.y) {
    const TI* pz = patches + (((int64_t)t * pd + cz + lz) * ph + cy) * (int64_t)pw * C + (int64_t)cx * C;
    TO* vz = out + (((int64_t)(oz + lz) * H + oy) * W + ox) * (int64_t)C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
      const int ly = i / row, r = i - ly * row;
      const float v = to_f<TI>(pz[(int64_t)ly * pw * C + r]);
      TO* o = vz + (int64_t)ly * W * C + r;
      *o = mode ? from_f<TO>(__fadd_rn(to_f<TO>(*o), v)) : from_f<TO>(v);
    }
  }
}

template <typename TI, typename TO>
static int launch_chunk_insert(const void* patches, void* out, const int64_t* desc, int64_t n, int H, int W, int C, int pd, int ph,
                               int pw, int mode, cudaStream_t st) {
  int64_t bx = ceil_div((int64_t)ph * pw * C, 256);
  if (bx > 32) bx = 32;
  dim3 blocks((unsigned)bx, (unsigned)pd, (unsigned)n);
  chunk_insert_kernel<TI, TO><<<blocks, 256, 0, st>>>((const TI*)patches, (TO*)out, desc, H, W, C, pd, ph, pw, mode);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace b200

B200_EXPORT int b200_chunk_extract(const void* src, int32_t dtype, int64_t D, int64_t H, int64_t W, int64_t C, void* dst, int64_t n,
                                   int64_t pd, int64_t ph, int64_t pw, const int64_t* desc, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(src && dst && desc, "chunk_extract: null pointer");
  B200_CHECK_ARG(valid_dtype(dtype) || dtype == 3, "chunk_extract: bad dtype");
  B200_CHECK_ARG(n > 0 && n <= 65535 && pd > 0 && pd <= 65535 && ph * pw * C < (1LL << 30), "chunk_extract: bad tile batch");
  B200_CHECK_ARG(D < (1LL << 31) && H < (1LL << 31) && W < (1LL << 31), "chunk_extract: volume axis too long");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t bx = ceil_div(ph * pw * C, 256);
  if (bx > 32) bx = 32;
  dim3 blocks((unsigned)bx, (unsigned)pd, (unsigned)n);
  if (dtype == 3)
    chunk_extract_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, desc, (int)H, (int)W, (int)C, (int)pd,
                                                          (int)ph, (int)pw);
  else if (dtype == B200_F32)
    chunk_extract_kernel<uint32_t><<<blocks, 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, desc, (int)H, (int)W, (int)C, (int)pd,
                                                           (int)ph, (int)pw);
  else
    chunk_extract_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)src, (uint16_t*)dst, desc, (int)H, (int)W, (int)C, (int)pd,
                                                           (int)ph, (int)pw);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_chunk_insert(const void* patches, int32_t dtype_in, int64_t n, int64_t pd, int64_t ph, int64_t pw, int64_t C,
                                  void* out, int32_t dtype_out, int64_t D, int64_t H, int64_t W, const int64_t* desc, int32_t mode,
                                  void* stream) {
  using namespace b200;
  B200_CHECK_ARG(patches && out && desc, "chunk_insert: null pointer");
  B200_CHECK_ARG(valid_dtype(dtype_in) && valid_dtype(dtype_out), "chunk_insert: bad dtype");
  B200_CHECK_ARG(mode == 0 || mode == 1, "chunk_insert: mode must be 0 (replace) or 1 (add)");
  B200_CHECK_ARG(n > 0 && n <= 65535 && pd > 0 && pd <= 65535 && ph * pw * C < (1LL << 30), "chunk_insert: bad tile batch");
  B200_CHECK_ARG(D < (1LL << 31) && H < (1LL << 31) && W < (1LL << 31), "chunk_insert: volume axis too long");
  cudaStream_t st = (cudaStream_t)stream;
#define CI(TI, TO) return launch_chunk_insert<TI, TO>(patches, out, desc, n, (int)H, (int)W, (int)C, (int)pd, (int)ph, (int)pw, mode, st)
  if (dtype_in == B200_F32 && dtype_out == B200_F32) CI(float, float);
  if (dtype_in == B200_F16 && dtype_out == B200_F16) CI(__half, __half);
  if (dtype_in == B200_F16 && dtype_out == B200_F32) CI(__half, float);
  if (dtype_in == B200_BF16 && dtype_out == B200_BF16) CI(__nv_bfloat16, __nv_bfloat16);
  if (dtype_in == B200_BF16 && dtype_out == B200_F32) CI(__nv_bfloat16, float);
  if (dtype_in == B200_F32 && dtype_out == B200_F16) CI(float, __half);
  if (dtype_in == B200_F32 && dtype_out == B200_BF16) CI(float, __nv_bfloat16);
#undef CI
  set_error("chunk_insert: unsupported dtype pair %d -> %d", dtype_in, dtype_out);
  return B200_ERR_UNSUPPORTED;
}
