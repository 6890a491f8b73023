// One launch for every weight pack / weight-gradient unpack of a training pass.  Each job is one of the element-wise
// permutations the product launches separately today (biapy_b200/csrc/conv_simt.cu: pack_weight_kernel, unpack_wgrad_kernel;
// conv_umma.cu: pack_weight_xfold_kernel, pack_convT_weight_kernel, unpack_convT_wgrad_kernel); a block looks up its job in a
// device-resident table by block index and walks the job's elements with the job's own block count.
#pragma once
#include <stdint.h>

enum PackKind : int32_t {
  PACK_PLAIN = 0,        // (Cout,Cin,taps) fp32 -> [Cout][tap][Cin] (flip: [Cin][flipped tap][Cout]) in T
  PACK_XFOLD = 1,        // block-Toeplitz packing of the x-folded kernels in T
  PACK_CONVT = 2,        // (Cin,Cout,taps) fp32 -> [tap][Cout][Cin] (flip = for_dgrad: [tap][Cin][Cout]) in T
  UNPACK_WGRAD = 3,      // [Cout][tap][Cin] fp32 -> (Cout,Cin,taps) fp32, flip = accumulate
  UNPACK_CONVT_WGRAD = 4, // [tap][Cout][Cin] fp32 -> (Cin,Cout,taps) fp32, flip = accumulate
  PACK_XLINE = 5         // rotated (48 x 16) SWIZZLE_32B tiles of the x-line kernel (conv_xline.cu: pack_weight_xline_kernel) in T
};

struct PackJob {
  const float* src;
  void* dst;
  int32_t kind;
  int32_t cout, cin;       // PACK_CONVT / UNPACK_CONVT_WGRAD: cout = Cin of the transposed conv, cin = its Cout (argument order
                           // of the product kernels: (cin, cout, taps))
  int32_t kd, kh, kw;      // plain / convT kinds use taps = kd * kh * kw
  int32_t flip;
  int32_t block_begin, n_blocks;
  int64_t total;           // elements of the job's index space
};

// window geometry of the x-folded packing (same rule as sm100::xfold_geom in conv_umma.cu)
__host__ __device__ inline bool pack_xfold_geom(int cin, int kw, int* xoff, int* kxp) {
  const int pw = kw / 2;
  if (cin % 16 == 0) {
    *xoff = 0;
    *kxp = (3 + kw) * cin;
    return *kxp % 32 == 0;
  }
  if (cin != 2 && cin != 4 && cin != 8) return false;
  int xo = 0;
  while (((pw + xo) * cin * 2) % 16 != 0) ++xo;
  *xoff = xo;
  *kxp = ((xo + 4 + 2 * pw) * cin + 31) / 32 * 32;
  return true;
}

template <typename T>
__device__ __forceinline__ void pack_batch_body(const PackJob* __restrict__ jobs, int n_jobs);

// job table in device memory (the emulator test drives this form)
template <typename T>
__global__ void __launch_bounds__(256) pack_batch_kernel(const PackJob* __restrict__ jobs, int n_jobs) {
  pack_batch_body<T>(jobs, n_jobs);
}

template <typename T>
__device__ __forceinline__ void pack_batch_body(const PackJob* __restrict__ jobs, int n_jobs) {
  __shared__ int s_job;
  if (threadIdx.x == 0) {
    int j = 0;
    while (j + 1 < n_jobs && (int)blockIdx.x >= jobs[j + 1].block_begin) ++j;
    s_job = j;
  }
  __syncthreads();
  const PackJob jb = jobs[s_job];
  const int64_t first = ((int64_t)blockIdx.x - jb.block_begin) * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)jb.n_blocks * blockDim.x;
  const int taps = jb.kd * jb.kh * jb.kw;
  if (jb.kind == PACK_PLAIN) {
    // walked by OUTPUT index: consecutive threads write consecutive elements (one 64-byte store per warp, not 32 scattered 2-byte
    // stores), and a block owns ONE contiguous chunk of the output, so the strided reads of a (co) / (ci) slab of the source stay
    // in L1 across its taps (27 * Cin * 4 bytes <= 41 KB)
    T* p = (T*)jb.dst;
    const uint32_t total = (uint32_t)jb.total, chunk = (total + jb.n_blocks - 1) / jb.n_blocks;
    const uint32_t lo = (uint32_t)(blockIdx.x - jb.block_begin) * chunk, hi = lo + chunk < total ? lo + chunk : total;
    const uint32_t cin = jb.cin, cout = jb.cout, tp = taps;
    for (uint32_t o = lo + threadIdx.x; o < hi; o += blockDim.x) {
      uint32_t co, ci, t;
      if (!jb.flip) { ci = o % cin; t = (o / cin) % tp; co = o / (cin * tp); }            // [Cout][tap][Cin]
      else { co = o % cout; t = tp - 1 - (o / cout) % tp; ci = o / (cout * tp); }         // [Cin][flipped tap][Cout]
      p[o] = from_f<T>(jb.src[((size_t)co * cin + ci) * tp + t]);
    }
  } else if (jb.kind == PACK_XFOLD) {
    T* out = (T*)jb.dst;
    const int CO = jb.flip ? jb.cin : jb.cout, CI = jb.flip ? jb.cout : jb.cin;
    int xoff = 0, kxp = 0;
    pack_xfold_geom(CI, jb.kw, &xoff, &kxp);
    const int win = kxp / CI;
    for (int64_t i = first; i < jb.total; i += stride) {
      int64_t t = i;
      const int ci = (int)(t % CI); t /= CI;
      const int xi = (int)(t % win); t /= win;
      const int dy = (int)(t % jb.kh); t /= jb.kh;
      const int dz = (int)(t % jb.kd); t /= jb.kd;
      const int co = (int)(t % CO);
      const int j = (int)(t / CO);
      const int kx = xi - xoff - j;
      float v = 0.f;
      if (kx >= 0 && kx < jb.kw) {
        if (!jb.flip) v = jb.src[((((int64_t)co * jb.cin + ci) * jb.kd + dz) * jb.kh + dy) * jb.kw + kx];
        else v = jb.src[((((int64_t)ci * jb.cin + co) * jb.kd + (jb.kd - 1 - dz)) * jb.kh + (jb.kh - 1 - dy)) * jb.kw + (jb.kw - 1 - kx)];
      }
      out[i] = from_f<T>(v);
    }
  } else if (jb.kind == PACK_CONVT) {
    T* p = (T*)jb.dst;
    const uint32_t cin = jb.cout, cout = jb.cin, tp = taps;          // see PackJob: product argument order (cin, cout, taps)
    const uint32_t total = (uint32_t)jb.total, chunk = (total + jb.n_blocks - 1) / jb.n_blocks;
    const uint32_t lo = (uint32_t)(blockIdx.x - jb.block_begin) * chunk, hi = lo + chunk < total ? lo + chunk : total;
    for (uint32_t o = lo + threadIdx.x; o < hi; o += blockDim.x) {   // by output index, as PACK_PLAIN
      uint32_t co, ci, t;
      if (!jb.flip) { ci = o % cin; co = (o / cin) % cout; t = o / (cin * cout); }        // [tap][Cout][Cin]
      else { co = o % cout; ci = (o / cout) % cin; t = o / (cout * cin); }                // [tap][Cin][Cout]
      p[o] = from_f<T>(jb.src[((size_t)ci * cout + co) * tp + t]);
    }
  } else if (jb.kind == PACK_XLINE) {
    T* out = (T*)jb.dst;
    const int CO = jb.flip ? jb.cin : jb.cout, CI = jb.flip ? jb.cout : jb.cin;
    const int ks = CI / 16;
    for (int64_t i = first; i < jb.total; i += stride) {
      // index space and layout of pack_weight_xline_kernel (conv_xline.cu): Cout' = 16: [r][dx][k][dy*48 + s*16 + co][kk];
      // Cout' = 48: [r][dy][dx][k][s*48 + co][kk]
      int64_t t = i;
      const int kk = (int)(t % 16); t /= 16;
      const int rr = (int)(t % 144); t /= 144;
      const int k = (int)(t % ks); t /= ks;
      const int dx = (int)(t % 3); t /= 3;
      int r, dy, s, co;
      if (CO == 16) { r = (int)t; dy = rr / 48; s = (rr % 48) / 16; co = rr % 16; }
      else { dy = (int)(t % 3); t /= 3; r = (int)t; s = rr / 48; co = rr % 48; }
      const int dz = ((r + 1 - s) % 3 + 3) % 3;
      const int ci = k * 16 + kk;
      float v;
      if (jb.flip) v = jb.src[((((int64_t)ci * jb.cin + co) * 3 + (2 - dz)) * 3 + (2 - dy)) * 3 + (2 - dx)];
      else v = jb.src[((((int64_t)co * jb.cin + ci) * 3 + dz) * 3 + dy) * 3 + dx];
      out[(i / (144 * 16)) * (144 * 16) + rr * 16 + ((((kk >> 3) ^ ((rr >> 2) & 1))) << 3) + (kk & 7)] = from_f<T>(v);
    }
  } else if (jb.kind == UNPACK_WGRAD) {
    float* dw = (float*)jb.dst;
    for (int64_t i = first; i < jb.total; i += stride) {
      const int t = (int)(i % taps);
      const int ci = (int)((i / taps) % jb.cin);
      const int co = (int)(i / ((int64_t)taps * jb.cin));
      const float v = jb.src[((int64_t)co * taps + t) * jb.cin + ci];
      dw[i] = jb.flip ? dw[i] + v : v;
    }
  } else {
    float* dw = (float*)jb.dst;
    const int cin = jb.cout, cout = jb.cin;
    for (int64_t i = first; i < jb.total; i += stride) {
      const int t = (int)(i % taps);
      const int co = (int)((i / taps) % cout);
      const int ci = (int)(i / ((int64_t)taps * cout));
      const float v = jb.src[((int64_t)t * cout + co) * cin + ci];
      dw[i] = jb.flip ? dw[i] + v : v;
    }
  }
}
