// Test-time augmentation by signed axis permutations (the 8 / 16 rotations + flips of biapy/data/post_processing/tta.py:65-260,
// driven by ensemble_predictions, biapy/data/post_processing/post_processing.py:1386-1555).  Two HBM-bound gathers:
//   orient_apply : dst = T(pad_front(src))           -- AxisTransform.apply (tta.py:158-166) fused with _pad_for_orientations
//   orient_reduce: out = crop(reduce_n T_n^-1(pred_n)) -- inverse transforms, _reduce_orientations and _crop_padding in one pass
// The reduction reproduces numpy's float32 arithmetic (sequential sum over the orientation axis, one true division), so the
// ensemble is bit-identical to the reference for float32 predictions.
#include "common.cuh"

namespace b200 {

struct OrientApplyParams {
  int perm[3], sign[3];     // output axis a comes from input axis perm[a], reversed when sign[a] < 0 (axes: z, y, x)
  int dim[3];               // source extents (unpadded)
  int pad[3];               // elements padded in FRONT of each source axis
  int odim[3];              // destination extents: odim[a] = dim[perm[a]] + pad[perm[a]]
  int mode;                 // 0 constant (zeros), 1 reflect, 3 edge  (the b200_pad_mode codes of the crop kernel)
  int n, c;
  int64_t sld, dld;
};

template <typename T>
__global__ void __launch_bounds__(256) orient_apply_kernel(const T* __restrict__ src, T* __restrict__ dst, const OrientApplyParams p) {
  const int64_t per = (int64_t)p.odim[0] * p.odim[1] * p.odim[2];
  const int64_t total = per * p.n * p.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % p.c);
    int64_t t = i / p.c;
    int v[3];
    v[2] = (int)(t % p.odim[2]); t /= p.odim[2];
    v[1] = (int)(t % p.odim[1]); t /= p.odim[1];
    v[0] = (int)(t % p.odim[0]); t /= p.odim[0];
    const int n = (int)t;
    int u[3];
    bool zero = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ax = p.perm[a];
      const int pos = p.sign[a] > 0 ? v[a] : p.odim[a] - 1 - v[a];       // coordinate in the padded source
      int s = pos - p.pad[ax];
      if (s < 0) {
        if (p.mode == 1) s = -s;                  // np.pad(..., 'reflect'): the edge element is not repeated
        else if (p.mode == 3) s = 0;              // 'edge'
        else zero = true;                         // 'constant'
      }
      u[ax] = s;
    }
    const int64_t so = ((((int64_t)n * p.dim[0] + u[0]) * p.dim[1] + u[1]) * p.dim[2] + u[2]) * p.sld + ch;
    const int64_t dof = ((((int64_t)n * p.odim[0] + v[0]) * p.odim[1] + v[1]) * p.odim[2] + v[2]) * p.dld + ch;
    dst[dof] = zero ? from_f<T>(0.f) : src[so];
  }
}

constexpr int kMaxOrient = 16;
struct OrientReduceParams {
  int perm[kMaxOrient][3], sign[kMaxOrient][3];
  int norient;
  int pdim[3];              // extents of every prediction (padded, square where axes are swapped)
  int pad[3];               // front padding to crop off
  int odim[3];              // output extents = pdim - pad
  int mode;                 // 0 mean, 1 min, 2 max
  int c;
  int64_t pld, old;
  int64_t pstride;          // elements between consecutive orientations
};

template <typename T>
__global__ void __launch_bounds__(256) orient_reduce_kernel(const T* __restrict__ pred, float* __restrict__ out, const OrientReduceParams p) {
  const int64_t total = (int64_t)p.odim[0] * p.odim[1] * p.odim[2] * p.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % p.c);
    int64_t t = i / p.c;
    int u[3];
    u[2] = (int)(t % p.odim[2]) + p.pad[2]; t /= p.odim[2];
    u[1] = (int)(t % p.odim[1]) + p.pad[1]; t /= p.odim[1];
    u[0] = (int)t + p.pad[0];
    float acc = 0.f;
    for (int k = 0; k < p.norient; ++k) {
      // the prediction of orientation k at v holds the voxel u of the (padded) image: v[a] = +-u[perm[a]]
      int v[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int s = u[p.perm[k][a]];
        v[a] = p.sign[k][a] > 0 ? s : p.pdim[a] - 1 - s;
      }
      const float val = to_f<T>(pred[(int64_t)k * p.pstride + (((int64_t)v[0] * p.pdim[1] + v[1]) * p.pdim[2] + v[2]) * p.pld + ch]);
      if (k == 0) acc = val;
      else if (p.mode == 0) acc = __fadd_rn(acc, val);
      else if (p.mode == 1) acc = fminf(acc, val);
      else acc = fmaxf(acc, val);
    }
    if (p.mode == 0) acc = __fdiv_rn(acc, (float)p.norient);
    out[i / p.c * p.old + ch] = acc;
  }
}

static bool valid_signed_perm(const int32_t* perm, const int32_t* sign) {
  int seen = 0;
  for (int a = 0; a < 3; ++a) {
    if (perm[a] < 0 || perm[a] > 2 || (sign[a] != 1 && sign[a] != -1)) return false;
    seen |= 1 << perm[a];
  }
  return seen == 7;
}

static int grid_for(int64_t total) {
  int64_t b = ceil_div(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace b200

B200_EXPORT int b200_orient_apply(const b200_tensor* src, const b200_tensor* dst, const int32_t* perm, const int32_t* sign,
                                  const int32_t* pad_before, int32_t pad_mode, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(check_tensor(src, "orient_apply.src") && check_tensor(dst, "orient_apply.dst"), "%s", b200_last_error());
  B200_CHECK_ARG(perm && sign && pad_before && valid_signed_perm(perm, sign), "orient_apply: (perm, sign) is not a signed axis permutation");
  B200_CHECK_ARG(src->dtype == dst->dtype && src->c == dst->c && src->n == dst->n, "orient_apply: src / dst dtype, channels and batch must match");
  B200_CHECK_ARG(pad_mode == 0 || pad_mode == 1 || pad_mode == 3, "orient_apply: pad_mode must be constant (0), reflect (1) or edge (3)");
  OrientApplyParams p{};
  const int sd[3] = {src->d, src->h, src->w}, dd[3] = {dst->d, dst->h, dst->w};
  for (int a = 0; a < 3; ++a) {
    p.perm[a] = perm[a]; p.sign[a] = sign[a]; p.dim[a] = sd[a]; p.pad[a] = pad_before[a];
    B200_CHECK_ARG(pad_before[a] >= 0, "orient_apply: negative padding");
    B200_CHECK_ARG(pad_mode != 1 || pad_before[a] < sd[a], "orient_apply: reflect padding needs pad < dim (np.pad rule; use edge)");
  }
  for (int a = 0; a < 3; ++a) {
    p.odim[a] = sd[perm[a]] + pad_before[perm[a]];
    B200_CHECK_ARG(dd[a] == p.odim[a], "orient_apply: dst axis %d is %d, expected %d", a, dd[a], p.odim[a]);
  }
  p.mode = pad_mode; p.n = src->n; p.c = src->c; p.sld = src->ld; p.dld = dst->ld;
  const int64_t total = (int64_t)p.n * p.odim[0] * p.odim[1] * p.odim[2] * p.c;
  B200_DISPATCH_DTYPE(src->dtype, T, {
    orient_apply_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)src->data, (T*)dst->data, p);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_orient_reduce(const b200_tensor* pred, const int32_t* perms, const int32_t* signs, int32_t mode,
                                   const int32_t* pad_before, const b200_tensor* out, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(check_tensor(pred, "orient_reduce.pred") && check_tensor(out, "orient_reduce.out"), "%s", b200_last_error());
  B200_CHECK_ARG(perms && signs && pad_before, "orient_reduce: null pointer");
  B200_CHECK_ARG(pred->n >= 1 && pred->n <= kMaxOrient, "orient_reduce: 1..%d orientations, got %d", kMaxOrient, pred->n);
  B200_CHECK_ARG(out->dtype == B200_F32 && out->n == 1 && out->c == pred->c, "orient_reduce: out must be float32, batch 1, same channels");
  B200_CHECK_ARG(mode >= 0 && mode <= 2, "orient_reduce: mode must be 0 (mean), 1 (min) or 2 (max)");
  OrientReduceParams p{};
  p.norient = pred->n;
  for (int k = 0; k < pred->n; ++k) {
    B200_CHECK_ARG(valid_signed_perm(perms + 3 * k, signs + 3 * k), "orient_reduce: orientation %d is not a signed axis permutation", k);
    for (int a = 0; a < 3; ++a) { p.perm[k][a] = perms[3 * k + a]; p.sign[k][a] = signs[3 * k + a]; }
  }
  const int pd[3] = {pred->d, pred->h, pred->w}, od[3] = {out->d, out->h, out->w};
  for (int a = 0; a < 3; ++a) {
    p.pdim[a] = pd[a]; p.pad[a] = pad_before[a]; p.odim[a] = od[a];
    B200_CHECK_ARG(pad_before[a] >= 0 && od[a] == pd[a] - pad_before[a], "orient_reduce: out axis %d is %d, expected %d", a, od[a],
                   pd[a] - pad_before[a]);
  }
  for (int k = 0; k < pred->n; ++k)
    for (int a = 0; a < 3; ++a)
      B200_CHECK_ARG(pd[a] == pd[p.perm[k][a]], "orient_reduce: orientation %d swaps axes of different length (%d vs %d)", k, pd[a],
                     pd[p.perm[k][a]]);
  p.mode = mode; p.c = pred->c; p.pld = pred->ld; p.old = out->ld;
  p.pstride = (int64_t)pred->d * pred->h * pred->w * pred->ld;
  const int64_t total = (int64_t)od[0] * od[1] * od[2] * p.c;
  B200_DISPATCH_DTYPE(pred->dtype, T, {
    orient_reduce_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)pred->data, (float*)out->data, p);
  });
  B200_LAUNCH_CHECK();
  return B200_OK;
}
