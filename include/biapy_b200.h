/*
 * biapy_b200 -- C ABI of the B200-native engine for BiaPy's patch U-Net hot path.
 *
 * This is the drop-in boundary: plain C, raw device pointers + sizes + cudaStream_t
 * (passed as void*), int status codes (0 = OK, <0 = error, text via b200_last_error()).  No allocation happens
 * inside any call: outputs and workspaces are caller-owned.  Calls are thread-compatible (no hidden global
 * state except the per-thread last-error string and a cache of TMA descriptors keyed by their arguments).
 *
 * The reference (BiaPy 3.7.0, 100% Python) has no FFI for this path: the "interface each entry point
 * replaces" is therefore the Python/torch call cited next to it (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a BiaPy maintainer would add.
 *
 * Tensors are channels-last: (N, D, H, W, C) with C contiguous -- the layout BiaPy's host arrays already
 * have (biapy/utils/misc.py:689-713 only permutes a (N,Z,Y,X,C) numpy array).  2D data uses D == 1.
 * `ld` is the voxel pitch in elements (>= C); it lets a tensor be a channel slice of a wider buffer so that
 * torch.cat([up, skip], 1) (biapy/models/blocks.py:664-666, 1653) is never materialised.
 */
#ifndef BIAPY_B200_H
#define BIAPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_ARG -1
#define B200_ERR_CUDA -2
#define B200_ERR_UNSUPPORTED -3

enum b200_dtype { B200_F32 = 0, B200_BF16 = 1, B200_F16 = 2, B200_U8 = 3 /* crop, image ends */, B200_U16 = 4 /* image ends only */ };
enum b200_act {
  B200_ACT_NONE = 0, B200_ACT_RELU = 1, B200_ACT_ELU = 2, B200_ACT_SILU = 3, B200_ACT_LEAKY_RELU = 4,
  B200_ACT_GELU = 5, B200_ACT_TANH = 6, B200_ACT_SIGMOID = 7, B200_ACT_SOFTPLUS = 8
};
enum b200_pad_mode { B200_PAD_ZEROS = 0, B200_PAD_REFLECT = 1, B200_PAD_SYMMETRIC = 2, B200_PAD_EDGE = 3, B200_PAD_WRAP = 4 };
enum b200_conv_impl { B200_IMPL_AUTO = 0, B200_IMPL_SIMT = 1, B200_IMPL_UMMA = 2, B200_IMPL_XFOLD = 3 };

typedef struct b200_tensor {
  void* data;      /* device pointer */
  int32_t dtype;   /* enum b200_dtype */
  int32_t n, d, h, w, c;
  int64_t ld;      /* voxel pitch in elements (>= c) */
} b200_tensor;

/* ------------------------------------------------------------------------------------------------ runtime */
const char* b200_last_error(void);
int b200_version(void);
/* sm count, compute capability major*10+minor, total global memory; any pointer may be NULL */
int b200_device_info(int device, int* sm_count, int* cc, int64_t* total_mem);

/* ------------------------------------------------------------------------------------- patch-grid planner
 * Bit-exact restatement (IEEE double + int64) of the per-axis grid arithmetic of
 *   crop_3D_data_with_overlap   biapy/data/data_3D_manipulation.py:536-563, 596-606  (frame = 0)
 *   merge_3D_data_with_overlap  biapy/data/data_3D_manipulation.py:777-816, 826-835  (frame = 1)
 * and their 2D twins biapy/data/data_2D_manipulation.py:226-243 / 464-484.  Host-only, no CUDA.          */
typedef struct b200_axis_plan {
  int64_t dim, patch, pad;   /* inputs */
  int64_t step, n, last;     /* step after the per-block overlap adjustment, patches on this axis, last-tile shift */
  int64_t core, ov_px;       /* patch - 2*pad; core - step (spline taper length of the merge) */
} b200_axis_plan;
/* returns B200_ERR_ARG with "division by zero" if int((patch-2*pad)*(1-overlap)) == 0 (the reference raises
 * ZeroDivisionError there) */
int b200_plan_axis(int64_t dim, int64_t patch, int64_t pad, double overlap, b200_axis_plan* out);
/* start of patch i: frame 0 = crop (padded frame), frame 1 = merge (original frame) */
int64_t b200_axis_start(const b200_axis_plan* plan, int64_t i, int frame);
/* 1-D spline taper, float32 (data_3D_manipulation.py:664-672): out[size] */
int b200_spline_window_1d(int64_t size, int64_t ov_px, float* out);

/* --------------------------------------------------------------------------------------- crop (device gather)
 * Replaces np.pad + the strided copies of crop_3D_data_with_overlap (data_3D_manipulation.py:505-516, 591-623).
 * src: one volume (D,H,W,C) dense, any of the three dtypes; dst: (n_patches, pd, ph, pw, C) dense, same dtype.
 * starts_{z,y,x}: per-axis patch starts in the padded frame (DEVICE int64 arrays of length n_{z,y,x});
 * patches are emitted in C order over (z,y,x), exactly the reference's `c` counter.                        */
int b200_crop_gather(const void* src, int32_t dtype, int64_t D, int64_t H, int64_t W, int64_t C,
                     void* dst, int64_t pd, int64_t ph, int64_t pw,
                     const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                     const int64_t* starts_x, int64_t nx,
                     int64_t pad_z, int64_t pad_y, int64_t pad_x, int32_t pad_mode, void* stream);

/* Only the patches [first, first + count) of the grid (C order over (z,y,x)); dst: (count, pd, ph, pw, C).  A rank of the
 * sharded sliding-window inference crops just the patches it predicts, and `src` may hold just the planes
 * [src_z0, src_z0 + src_nz) of the D-plane volume (what that rank read / uploaded). */
int b200_crop_gather_range(const void* src, int32_t dtype, int64_t D, int64_t H, int64_t W, int64_t C,
                           void* dst, int64_t pd, int64_t ph, int64_t pw,
                           const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                           const int64_t* starts_x, int64_t nx,
                           int64_t pad_z, int64_t pad_y, int64_t pad_x, int32_t pad_mode, int64_t first, int64_t count,
                           int64_t src_z0, int64_t src_nz, void* stream);

/* ------------------------------------------------------------------------ overlap-add merge (device gather)
 * Replaces the accumulate + normalise loop of merge_3D_data_with_overlap (data_3D_manipulation.py:822-849).
 * Gather formulation: one thread per output element walks the covering patches in increasing patch index
 * and reproduces the reference's float32 operation order (multiply, add, add, divide) -- results are
 * bit-identical to numpy for float32 / float16 patches.
 * patches: (n_patches, pz, py, px, C) dense INCLUDING the `pad` border (it is skipped, :748-755);
 * out: (D,H,W,C) dense, dtype_out (f32 / bf16 / f16; the reference casts to the patch dtype).
 * win_{z,y,x}: device float32 1-D windows of the core size; starts_*: device int64 merge starts (orig frame). */
int b200_overlap_add(const void* patches, int32_t dtype_in, void* out, int32_t dtype_out,
                     int64_t D, int64_t H, int64_t W, int64_t C,
                     int64_t pz, int64_t py, int64_t px, int64_t pad_z, int64_t pad_y, int64_t pad_x,
                     const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                     const int64_t* starts_x, int64_t nx,
                     const float* win_z, const float* win_y, const float* win_x, void* stream);

/* The same merge restricted to the output planes [z0, z0 + nz_out): `out` is (nz_out, H, W, C) dense.  `patches` is still the
 * full (n_patches, ...) array, of which only the elements covering the slab are read -- sharded sliding-window inference lets
 * every rank own one z slab of the volume and receive just those pieces of its neighbours' patch predictions. */
int b200_overlap_add_slab(const void* patches, int32_t dtype_in, void* out, int32_t dtype_out,
                          int64_t D, int64_t H, int64_t W, int64_t C,
                          int64_t pz, int64_t py, int64_t px, int64_t pad_z, int64_t pad_y, int64_t pad_x,
                          const int64_t* starts_z, int64_t nz, const int64_t* starts_y, int64_t ny,
                          const int64_t* starts_x, int64_t nx,
                          const float* win_z, const float* win_y, const float* win_x, int64_t z0, int64_t nz_out, void* stream);

/* ------------------------------------------------------------------------------ by-chunks tile grid (host)
 * The reference's multi-GPU inference path (TEST.BY_CHUNKS): non-blended tiles of (crop - 2*pad), read with a halo,
 * reflect-padded to the crop shape, written back without the halo.  Integer bookkeeping, bit-exact with
 *   chunked_test_pair_data_generator.__init__   biapy/data/generators/chunked_test_pair_data_generator.py:276-295
 *   chunked_test_pair_data_generator._patch_coords                                            ...:440-486
 *   extract_and_prepare_sample (pad_to_add and the "real padding info")                       ...:536-560
 * Host-only, no CUDA.                                                                                        */
typedef struct b200_chunk_grid {
  int64_t dim[3], crop[3], pad[3];           /* inputs (z, y, x) */
  int64_t step[3], vols[3];                  /* crop - 2*pad; ceil(dim / step) */
  int64_t z_vol_start, z_vol_end, total;     /* tile range of the requested Z slab; tiles to process */
} b200_chunk_grid;
/* z_start / z_end = -1: whole volume.  Errors (B200_ERR_ARG) mirror the reference's ValueErrors (:247-274). */
int b200_chunk_grid_plan(const int64_t dim[3], const int64_t crop[3], const int64_t pad[3], int64_t z_start, int64_t z_end,
                         b200_chunk_grid* out);
/* out[27] for tile `vol_id`: grid position z,y,x | region to read [zs,ze,ys,ye,xs,xe] (halo included, clipped) |
 * region to write back [zs,ze,ys,ye,xs,xe] | np.pad amounts [zl,zr,yl,yr,xl,xr] | real padding info (max(pad, amount)) */
int b200_chunk_patch_coords(const b200_chunk_grid* grid, int64_t vol_id, int64_t out[27]);

/* -------------------------------------------------------------------------- by-chunks extract / insert (device)
 * b200_chunk_extract replaces extract_patch_within_image + np.pad(..., "reflect") (chunked_test_pair_data_generator.py
 * :505-551): tile t of `n` is the window [start, start+len) per axis of the resident volume (D,H,W,C), reflected about the
 * WINDOW's own ends (numpy pads the extracted array, not the volume) and shifted by `left`.
 *   desc (DEVICE int64 [n][9]) = start_z, len_z, left_z, start_y, len_y, left_y, start_x, len_x, left_x
 * dst: (n, pd, ph, pw, C) dense, same dtype as src (f32 / bf16 / f16 / u8).
 * b200_chunk_insert replaces the strip (base_workflow.py:2606-2612) + insert_patch_in_efficient_file
 * (data_3D_manipulation.py:286-351; mode 0 = "replace", 1 = "add"): the region [strip, strip+size) of tile t goes to
 * out[o : o+size).   desc (DEVICE int64 [n][9]) = out_z, size_z, strip_z, out_y, size_y, strip_y, out_x, size_x, strip_x
 * patches: (n, pd, ph, pw, C) dense; out: (D,H,W,C) dense; dtype pairs as in b200_overlap_add.               */
int b200_chunk_extract(const void* src, int32_t dtype, int64_t D, int64_t H, int64_t W, int64_t C, void* dst, int64_t n,
                       int64_t pd, int64_t ph, int64_t pw, const int64_t* desc, void* stream);
int b200_chunk_insert(const void* patches, int32_t dtype_in, int64_t n, int64_t pd, int64_t ph, int64_t pw, int64_t C,
                      void* out, int32_t dtype_out, int64_t D, int64_t H, int64_t W, const int64_t* desc, int32_t mode,
                      void* stream);

/* b200_select_hist: one pass of a radix select over channel `ch` of an interleaved image (percentile clipping, norm.py:445-466:
 * np.percentile / torch kthvalue need exact order statistics).  Values map to 32-bit keys that sort like the values (unsigned
 * integers unchanged; float32 with the sign bit flipped / negatives inverted); hist[d] (device, 1 << bits counters, cleared by the
 * call) = number of elements whose key has digit d = (key >> shift) & (2^bits - 1) and, when has_prefix, whose higher bits
 * key >> (shift + bits) equal `prefix`.  bits <= 11. */
int b200_select_hist(const void* src, int32_t dtype, int64_t voxels, int32_t c, int32_t ch, int32_t shift, int32_t bits,
                     uint32_t prefix, int32_t has_prefix, uint32_t* hist /* device */, void* stream);
/* b200_edge_hist: counts (device uint64[nbins], cleared by the call) = np.histogram(src, bins=nbins)[0] for a float32 array and the
 * bin edges numpy derives from its [min, max] (device float32[nbins + 1], computed by the caller with np.linspace): numpy's
 * equal-bin index rule incl. its edge corrections.  The heavy half of skimage.filters.threshold_otsu (scikit-image >= 0.21,
 * pyproject.toml:26), which after_merge_patches / after_full_image call on the merged prediction (semantic_seg.py:429, 455). */
int b200_edge_hist(const float* src, int64_t n, const float* edges, int32_t nbins, uint64_t* counts, void* stream);
/* ------------------------------------------------------------------------------------- the ends of the path
 * Image normalisation in front of the first convolution and its inverse / the binarisation behind the merge.
 * Images are dense (voxels, C) arrays of dtype u8 / u16 / f32.
 * b200_image_stats: one pass, per channel c: out[c] = (min, max, sum, sum of squares, is_binary) as doubles (is_binary = 1 when
 *   every value is 0 or 1: biapy/data/norm.py:38-42 -- binary channels are never normalised).  clip (host, [c][3] = on, lo, hi,
 *   or NULL): statistics of min(max(x, lo), hi), what the reference sees after its in-place percentile clip.
 * b200_image_norm_apply = the per-channel body of normalize_image (norm.py:188-218) in float32, same operation order:
 *   params[c] = (clip, lo, hi, kind, a, b):  clip != 0 -> x = min(max(x, lo), hi)              (percentile_clip :468-472)
 *     kind 1 -> x = (x - a) / b   a = min_val_to_div, b = max(max_val_to_div - min_val_to_div, eps)      (norm_range01 :575-578)
 *     kind 2 -> x = (x - a) / b   a = mean,           b = max(std, eps)       (zero_mean_unit_variance_normalization :632-633)
 *     kind 0 -> x unchanged (binary channel).   dst: float32.
 * b200_image_denorm_apply = undo_image_norm (norm.py:641-780) in float64 like numpy (float32 data * Python-list factors):
 *   kind 1: clip(x, 0, 1) * a + b  (a = max_val_to_div, b = min_val_to_div, :703-713);  kind 2: x * a + b (a = std, b = mean),
 *   rounded and clamped to the integer range when dst is u8 / u16 (:767-778); conversion to dst dtype as numpy's astype
 *   (truncation for integers).  src: float32; dst: u8 / u16 / f32.
 * b200_binarize: dst u8 = src > threshold (semantic_seg.py:420-421, by-chunks fixed 0.5 :524-529);
 * b200_argmax_channels: dst (u8 / u16, 1 channel) = first index of the channel maximum (np.argmax, :423-424, :531).      */
int b200_image_stats(const void* src, int32_t dtype, int64_t voxels, int32_t c, const float* clip, double* out /* [c][5], device */,
                     void* stream);
int b200_image_norm_apply(const void* src, int32_t dtype, int64_t voxels, int32_t c, const float* params /* [c][6], host */,
                          float* dst, void* stream);
int b200_image_denorm_apply(const float* src, int64_t voxels, int32_t c, const double* params /* [c][3] kind,a,b, host */,
                            void* dst, int32_t dst_dtype, void* stream);
int b200_binarize(const float* src, int64_t n, float threshold, uint8_t* dst, void* stream);
int b200_argmax_channels(const float* src, int64_t voxels, int32_t c, void* dst, int32_t dst_dtype, void* stream);

/* ---------------------------------------------------------------------------------- test-time augmentation
 * Signed axis permutations of biapy/data/post_processing/tta.py:65-260 (AxisTransform: output axis a comes from input axis
 * perm[a], reversed when sign[a] == -1; axes are (z, y, x), 2D uses D = 1 with perm[0] = 0) as used by
 * ensemble_predictions (biapy/data/post_processing/post_processing.py:1386-1555) for scalar predictions (tta_spec = None).
 * b200_orient_apply  = AxisTransform.apply (tta.py:158-166) on every batch element of src, fused with the front padding of
 *   _pad_for_orientations (post_processing.py:1285-1339): dst dims = (src dims + pad_before) permuted; pad_mode as in
 *   b200_crop_gather (0 constant, 1 reflect, 3 edge).  src, dst: same dtype.
 * b200_orient_reduce = steps 4-5 of ensemble_predictions: pred holds one prediction per orientation (batch axis = orientation,
 *   perms / signs: [N][3], the FORWARD transforms), each is read through its inverse, reduced with mode 0 mean / 1 min / 2 max
 *   (_reduce_orientations :1349-1383; the mean adds the orientations in order in float32 and divides once, like np.mean)
 *   and the front padding is cropped (_crop_padding :1342-1346).  out: float32, batch 1, dims = pred dims - pad_before.    */
int b200_orient_apply(const b200_tensor* src, const b200_tensor* dst, const int32_t* perm, const int32_t* sign,
                      const int32_t* pad_before, int32_t pad_mode, void* stream);
int b200_orient_reduce(const b200_tensor* pred, const int32_t* perms, const int32_t* signs, int32_t mode,
                       const int32_t* pad_before, const b200_tensor* out, void* stream);

/* ---------------------------------------------------------------------------------------------- convolution
 * nn.Conv3d / nn.Conv2d, stride 1, padding='same', bias (biapy/models/blocks.py:154,157,1372; unet.py:347).
 * Weights are packed once per step from the PyTorch layout (Cout,Cin,kd,kh,kw) fp32:
 *   flip_transpose = 0 : [Cout][tap][Cin]          (fprop operand, K = tap*Cin + ci contiguous)
 *   flip_transpose = 1 : [Cin][flipped tap][Cout]  (dgrad operand: dX = conv(dY, W'))                     */
int b200_pack_conv_weight(const float* w, void* packed, int32_t dtype, int32_t cout, int32_t cin,
                          int32_t kd, int32_t kh, int32_t kw, int32_t flip_transpose, void* stream);
/* y = conv(x, w) + bias (+ residual) ; if accumulate != 0: y += ... (used by dgrad into a shared gradient) */
int b200_conv_fprop(const b200_tensor* x, const void* w_packed, const float* bias, const b200_tensor* residual,
                    const b200_tensor* y, int32_t kd, int32_t kh, int32_t kw, int32_t accumulate, int32_t impl,
                    void* stream);
/* dw_packed[Cout][tap][Cin] (fp32) += sum_vox dy[vox][co] * x[vox+tap][ci] ; dbias[Cout] += sum_vox dy.
 * Both outputs must be zero-initialised by the caller (they are accumulated with atomics).                  */
int b200_conv_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dw_packed, float* dbias,
                    int32_t kd, int32_t kh, int32_t kw, int32_t impl, void* stream);
/* Convolution (x-folded kernel family, weights from b200_pack_conv_weight_xfold) with the channel statistics of the
 * following normalisation fused into the epilogue -- the "GroupNorm statistics in the producer epilogue" half of the
 * Conv3D -> GN -> SiLU fusion (reference order blocks.py:154-160):
 *   sums[n][c] += (sum y, sum y*y) over the stored (rounded) output            == b200_channel_sums(y)
 * sums: double [N][C][2], zero-initialised by the caller.  *applied = 1 if the statistics were produced inside the
 * convolution, 0 if this shape has no fused epilogue (Cout != 16) or the call accumulates (the convolution still ran; call
 * b200_channel_sums).                                                                                                   */
int b200_conv_fprop_stats(const b200_tensor* x, const void* w_packed_xfold, const float* bias, const b200_tensor* y,
                          int32_t kd, int32_t kh, int32_t kw, int32_t accumulate, double* sums, int32_t* applied, void* stream);
/* x-folded tensor-core kernel for small-channel layers with kw in {1, 3} (see csrc/conv_umma.cu): block-Toeplitz packing
 *   [(j, co)][(dz, dy, xi, ci)], j < 4, xi < 3+kw; elements = 4*Cout' * kd*kh*(3+kw)*Cin'
 * where (Cout', Cin') = (Cout, Cin), or (Cin, Cout) when flip_transpose (dgrad operand).  Used with impl =
 * B200_IMPL_XFOLD in b200_conv_fprop; B200_IMPL_AUTO never selects it (the packing differs). */
int b200_pack_conv_weight_xfold(const float* w, void* packed, int32_t dtype, int32_t cout, int32_t cin, int32_t kd,
                                int32_t kh, int32_t kw, int32_t flip_transpose, void* stream);
/* x-line kernel (csrc/conv_xline.cu): 3x3x3 convolution of the Cout = 16 layers at W = 128 (Cin = 16 or 48, dense channels-last
 * input) -- one GEMM row per voxel, the A operand (a line and its two x-shifted copies) in tensor memory, no block-Toeplitz zeros
 * -- and the "Conv3D + GroupNorm + SiLU" fusion of the reference's pre-activation order GN(in) -> act -> conv
 * (biapy/models/blocks.py:1304-1378; conv -> norm -> act of blocks.py:148-160 is this launch's `sums` + the next one's `fuse`):
 *   fuse = 0: y (+)= conv(x) + bias;
 *   fuse = 1 / 2: a = silu(x * scale[n,c] + shift[n,c]) -- the arithmetic of b200_scale_shift_act (1) / b200_scale_shift_silu_fast
 *     (2), rounded to the engine dtype exactly like their stored result -- then y (+)= conv(a) + bias; a_out (nullable; dense, x's
 *     shape) receives a for the backward pass, bit-identical to the stand-alone launch.
 * sums (nullable): double [N][16][2] += (sum y, sum y*y) of the stored output, as b200_conv_fprop_stats.  accumulate on dense
 * output lines without sums is an element-wise add of the bulk-copy engine (formed in L2 on the rounded value).
 * Weights: b200_pack_conv_weight_xline -- 9 * Cin/16 tiles [r][dx][k] of 144 x 16 elements, row dy*48 + s*16 + co, column ci - 16k,
 * tap dz = (r + 1 - s) mod 3 (r = input plane mod 3, s = output plane mod 3), each tile in the SWIZZLE_32B K-major shared-memory
 * image; flip_transpose = 1 packs the dgrad operand.  b200_conv_xline_supported: 1 when (x, y, kernel) qualify and B200_XLINE != 0. */
int b200_conv_xline_supported(const b200_tensor* x, const b200_tensor* y, int32_t kd, int32_t kh, int32_t kw);
int b200_pack_conv_weight_xline(const float* w, void* packed, int32_t dtype, int32_t cout, int32_t cin, int32_t flip_transpose,
                                void* stream);
int b200_conv_fprop_xline(const b200_tensor* x, const void* w_packed, const float* bias, const b200_tensor* y, int32_t accumulate,
                          const float* scale, const float* shift, int32_t fuse, const b200_tensor* a_out, double* sums,
                          void* stream);
/* x-line weight gradient (csrc/conv_xline.cu): dw_packed[16][27][Cin] (fp32, zero-initialised or accumulated by the caller) +=
 * sum_vox dy[vox][co] * x[vox + off(tap)][ci] for the 3x3x3 layers with 16 output channels at W = 128, Cin = 16 or 48, dense
 * 16-bit lines; dbias (nullable) += sum_vox dy (inside the same launch: a row of ones in the activation operand).  The contraction
 * runs over the 128 voxels of a line with the transposed
 * activation line in tensor memory and the transposed dY lines as the shared-memory operand; the 27-tap gradient block stays in
 * tensor memory for the whole launch (one launch per 16 input channels).  Same contract as b200_conv_wgrad. */
int b200_conv_wgrad_xline_supported(const b200_tensor* x, const b200_tensor* dy, int32_t kd, int32_t kh, int32_t kw);
int b200_conv_wgrad_xline(const b200_tensor* x, const b200_tensor* dy, float* dw_packed, float* dbias, void* stream);
/* one tcgen05.mma with the A operand in tensor memory against exact integers; *max_err = largest absolute deviation */
int b200_xline_selftest(double* max_err, int32_t verbose, void* stream);
/* best kernel family for these operands: B200_IMPL_XFOLD, B200_IMPL_UMMA or B200_IMPL_SIMT
 * (wgrad != 0: for b200_conv_wgrad with y = dy; never XFOLD) */
int b200_conv_impl_query(const b200_tensor* x, const b200_tensor* y, int32_t kd, int32_t kh, int32_t kw, int32_t wgrad);
/* dw (Cout,Cin,kd,kh,kw) fp32 (+)= dw_packed[Cout][tap][Cin] */
int b200_unpack_conv_wgrad(const float* dw_packed, float* dw, int32_t cout, int32_t cin, int32_t taps,
                           int32_t accumulate, void* stream);

/* -------------------------------------------------------------------------- transposed convolution (k == s)
 * nn.ConvTranspose3d(kernel_size=s, stride=s) (biapy/models/blocks.py:603, 1607).  w: PyTorch layout
 * (Cin,Cout,sd,sh,sw) fp32, read directly.                                                                   */
int b200_convT_fprop(const b200_tensor* x, const float* w, const float* bias, const b200_tensor* y,
                     int32_t sd, int32_t sh, int32_t sw, void* stream);
int b200_convT_dgrad(const b200_tensor* dy, const float* w, const b200_tensor* dx,
                     int32_t sd, int32_t sh, int32_t sw, int32_t accumulate, void* stream);
/* dw (Cin,Cout,sd,sh,sw), dbias (Cout): zero-initialised by the caller, accumulated with atomics */
int b200_convT_wgrad(const b200_tensor* x, const b200_tensor* dy, float* dw, float* dbias,
                     int32_t sd, int32_t sh, int32_t sw, void* stream);

/* Tensor-core route of the same three passes (16-bit dtypes, Cin/Cout multiples of 16): each of the s^3 output phases
 * is a pointwise tcgen05 GEMM over a strided sub-lattice view of the fine tensor.  Weights are packed per step:
 *   for_dgrad = 0: [tap][Cout][Cin] (fprop)        for_dgrad = 1: [tap][Cin][Cout] (dgrad)
 * dw_packed of the wgrad is fp32 [tap][Cout][Cin], zero-initialised by the caller.                                   */
int b200_convT_tc_supported(const b200_tensor* x, const b200_tensor* y, int32_t sd, int32_t sh, int32_t sw);
int b200_pack_convT_weight(const float* w, void* packed, int32_t dtype, int32_t cin, int32_t cout, int32_t taps,
                           int32_t for_dgrad, void* stream);
int b200_convT_fprop_tc(const b200_tensor* x, const void* w_packed, const float* bias, const b200_tensor* y,
                        int32_t sd, int32_t sh, int32_t sw, void* stream);
int b200_convT_dgrad_tc(const b200_tensor* dy, const void* w_packed_t, const b200_tensor* dx,
                        int32_t sd, int32_t sh, int32_t sw, int32_t accumulate, void* stream);
int b200_convT_wgrad_tc(const b200_tensor* x, const b200_tensor* dy, float* dw_packed, float* dbias,
                        int32_t sd, int32_t sh, int32_t sw, void* stream);
int b200_unpack_convT_wgrad(const float* dw_packed, float* dw, int32_t cin, int32_t cout, int32_t taps,
                            int32_t accumulate, void* stream);

/* -------------------------------------------------------------------------------------------------- pooling
 * nn.MaxPool3d / MaxPool2d with kernel == stride (unet.py:255-256).  Backward recomputes the arg-max from
 * (x, y): the first maximum in (d,h,w) scan order receives the gradient, as ATen's max_pool backward.       */
int b200_maxpool_fwd(const b200_tensor* x, const b200_tensor* y, int32_t pd, int32_t ph, int32_t pw, void* stream);
int b200_maxpool_bwd(const b200_tensor* x, const b200_tensor* y, const b200_tensor* dy, const b200_tensor* dx,
                     int32_t pd, int32_t ph, int32_t pw, int32_t accumulate, void* stream);
/* Out-of-place form for gradients that live in a channel slice of a concat buffer: dx_out = (dx_in ? dx_in : 0) + the routed dy,
 * dx_out dense.  The max-pool backward is the last writer of a skip tensor's gradient, so handing the sum on as a dense tensor
 * saves the strided-to-dense copy the x-folded dgrad / wgrad of the producing block would need.  2 x 2 (x 2) windows, 16-bit. */
int b200_maxpool_bwd_to_ok(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx_in, const b200_tensor* dx_out,
                           int32_t pd, int32_t ph, int32_t pw);
int b200_maxpool_bwd_to(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx_in, const b200_tensor* dx_out,
                        int32_t pd, int32_t ph, int32_t pw, void* stream);

/* ------------------------------------------------------------------------------------ normalisation + activation
 * GroupNorm(8|16, C) / InstanceNorm(affine) / BatchNorm (biapy/models/blocks.py:2092-2165) as
 * (1) per-(n,c) sums, (2) finalize to per-(n,c) scale/shift, (3) fused scale-shift-activation.
 * sums: double [N][C][2] = (sum x, sum x^2), zero-initialised by the caller.                                  */
int b200_channel_sums(const b200_tensor* x, double* sums, void* stream);
/* groups: G channels groups per sample (G == C -> instance norm). batch_stats != 0: statistics are pooled over
 * the batch as well (BatchNorm training).  Writes mean[N][G], rstd[N][G], scale[N][C], shift[N][C] (fp32).   */
int b200_norm_finalize(const double* sums, int32_t n, int32_t c, int32_t groups, int64_t spatial, int32_t batch_stats,
                       const float* gamma, const float* beta, float eps,
                       float* mean, float* rstd, float* scale, float* shift, void* stream);
/* nn.Dropout(p), training mode (reference blocks.py:162-163): y (+)= keep ? x/(1-p) : 0 with a counter-based mask that is
 * a function of (*seed_dev, layer_id, element index); calling it on dy with the same (seed, layer_id) is the backward.   */
int b200_dropout(const b200_tensor* x, const b200_tensor* y, float p, const int64_t* seed_dev, int64_t layer_id,
                 int32_t accumulate, void* stream);
/* nn.Upsample(mode = 'bilinear' | 'trilinear', align_corners = False) with integer scale factors = y dims / x dims
 * (reference blocks.py:604-606, the 'upsampling' decoder variant); bwd is its adjoint in gather form.            */
int b200_upsample_linear_fwd(const b200_tensor* x, const b200_tensor* y, void* stream);
int b200_upsample_linear_bwd(const b200_tensor* dy, const b200_tensor* dx, int32_t accumulate, void* stream);
/* BatchNorm bookkeeping (torch.nn.BatchNorm{2,3}d / SyncBatchNorm semantics, reference blocks.py:2117-2120, 2155-2158):
 * running_mean/var <- (1-momentum)*running + momentum*(batch mean, UNBIASED batch variance); `mean`/`rstd` are the
 * first C entries written by b200_norm_finalize(groups = C, batch_stats = 1), `count` = N*D*H*W (x world size).   */
int b200_bn_update_running(const float* mean, const float* rstd, float eps, double count, float momentum,
                           float* running_mean, float* running_var, int32_t c, void* stream);
/* eval mode: scale[n,c] = gamma/sqrt(running_var+eps), shift[n,c] = beta - running_mean*scale (for b200_scale_shift_act) */
int b200_bn_eval_coeffs(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                        float eps, int32_t n, int32_t c, float* scale, float* shift, void* stream);
/* y = act(x * scale[n,c] + shift[n,c]);  scale/shift may be NULL (pure activation)                           */
int b200_scale_shift_act(const b200_tensor* x, const float* scale, const float* shift, int32_t act,
                         const b200_tensor* y, void* stream);
/* backward, pass 1: with xhat = (x-mean)*rstd, ypre = xhat*gamma+beta, g = dy*act'(ypre):
 *   red[N][C][2] (double, zero-initialised) += (sum g, sum g*(x - mean))   [centred: sum g*xhat = rstd * it]     */
int b200_norm_act_bwd_reduce(const b200_tensor* x, const b200_tensor* dy, const float* mean, const float* rstd,
                             int32_t groups, const float* gamma, const float* beta, int32_t act,
                             double* red, void* stream);
/* tiny: coef[N][C][4] = (k0, B, P, Q) such that, with ypre = x*k0 + B and g = dy*act'(ypre), dx = g*k0 - x*P - Q;
 * dgamma[C] += sum_n S2, dbeta[C] += sum_n S1                                                                  */
int b200_norm_bwd_finalize(const double* red, const float* mean, const float* rstd, const float* gamma, const float* beta,
                           int32_t n, int32_t c, int32_t groups, int64_t spatial, int32_t batch_stats,
                           float* coef, float* dgamma, float* dbeta, void* stream);
/* the same, and dxsum[C] += sum over samples and voxels of the dx the apply pass will write (not accumulated), from the sums alone:
 * xsums[N][C][2] are the forward channel sums (b200_channel_sums / the convolution epilogue).  That is the bias gradient of a
 * convolution whose output feeds this normalisation only (reference: autograd of blocks.py:148-160) -- no pass over dx.          */
int b200_norm_bwd_finalize_sums(const double* red, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                int32_t n, int32_t c, int32_t groups, int64_t spatial, int32_t batch_stats,
                                float* coef, float* dgamma, float* dbeta, const double* xsums, float* dxsum, void* stream);
/* xsum[Cin] += W^T dysum for a pointwise convolution W (Cout, Cin): channel sums of a gradient after a 1x1 layer's input gradient
 * has been accumulated into it                                                                                                    */
int b200_sums_through_pointwise(const float* w, const float* dysum, float* xsum, int32_t cout, int32_t cin, void* stream);
/* pass 2: dx (+)= g*k0 - x*P - Q                                                                               */
int b200_norm_act_bwd_apply(const b200_tensor* x, const b200_tensor* dy, int32_t act, const float* coef,
                            const b200_tensor* dx, int32_t accumulate, void* stream);
/* GroupNorm / InstanceNorm + SiLU chain of the 16-bit engine with cheaper arithmetic (one MUFU.TANH per sigmoid, activation
 * derivative evaluated once).  Same mathematics as scale_shift_act / norm_act_bwd_reduce / norm_act_bwd_apply for act = SiLU
 * (reference blocks.py:148-160 norm -> act), different contract: b200_norm_silu_bwd_reduce_g OVERWRITES dy with
 * g = dy * silu'(norm(x)) and b200_norm_bwd_apply_g consumes that g.  b200_norm_silu_fast_ok: 1 when (x, dy, dx) qualify
 * (bf16 / fp16, 8-channel vectors, <= 2048 channels); dy / dx may be null for the forward-only question. */
int b200_norm_silu_fast_ok(const b200_tensor* x, const b200_tensor* dy, const b200_tensor* dx);
int b200_scale_shift_silu_fast(const b200_tensor* x, const float* scale, const float* shift, const b200_tensor* y, void* stream);
/* write_g = 0: dy is left alone and the apply pass is b200_norm_silu_bwd_apply_fast, which recomputes the derivative (no extra
 * write in pass 1, one more MUFU per element in pass 2). */
int b200_norm_silu_bwd_reduce_g(const b200_tensor* x, const b200_tensor* dy_g, const float* mean, const float* rstd,
                                int32_t groups, const float* gamma, const float* beta, double* red, int32_t write_g, void* stream);
int b200_norm_silu_bwd_apply_fast(const b200_tensor* x, const b200_tensor* dy, const float* coef, const b200_tensor* dx,
                                  int32_t accumulate, void* stream);
int b200_norm_bwd_apply_g(const b200_tensor* x, const b200_tensor* g, const float* coef, const b200_tensor* dx,
                          int32_t accumulate, void* stream);
/* activation only (norm == 'none'): dx (+)= dy * act'(x) */
int b200_act_bwd(const b200_tensor* x, const b200_tensor* dy, int32_t act, const b200_tensor* dx,
                 int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------- element-wise
 * op: 0 y=a+b, 1 y=a*b, 2 y=relu(a+b), 3 y=copy(a), 4 y=sigmoid(a); b may have c == 1 (broadcast over channels) */
int b200_binary(const b200_tensor* a, const b200_tensor* b, const b200_tensor* y, int32_t op, void* stream);
/* attention gate backward helpers (biapy/models/blocks.py:1112-1116):
 *   out = psi * x : dpsi[vox] = sum_c dout*x ; dx (+)= dout*psi                                                */
int b200_gate_bwd(const b200_tensor* x, const b200_tensor* psi, const b200_tensor* dout,
                  const b200_tensor* dpsi, const b200_tensor* dx, int32_t accumulate, void* stream);
/* dy -> da: (a+b) relu'd : da = dy * (y > 0) */
int b200_relu_mask_bwd(const b200_tensor* y, const b200_tensor* dy, const b200_tensor* da, void* stream);
/* dtype conversion / strided copy between tensor views */
int b200_convert(const b200_tensor* src, const b200_tensor* dst, void* stream);

/* ------------------------------------------------------------------------------------------------- losses
 * BCEWithLogits mean (biapy/engine/metrics.py:544-546, 577-580): loss_sum[0] (double, zeroed by caller) +=
 * sum of per-element losses; dlogits = (sigmoid(z) - t) * grad_scale.  target is float32 dense.                */
int b200_bce_logits(const b200_tensor* logits, const float* target, double* loss_sum, const b200_tensor* dlogits,
                    float grad_scale, void* stream);
/* N2V masked MSE (metrics.py:2265-2286).  mode 0: sums[0] += sum((t - y*m)^2), sums[1] += sum(m); mode 1:
 * dpred = -2*(t - y*m)*m * grad_scale (grad_scale = upstream / sum(m)); mode 2: both in one pass (the division by sum(m)
 * is then left to b200_optim_step_dev's `denom`).  target dense (N,...,2C) float32.                                */
int b200_n2v_mse(const b200_tensor* pred, const float* target, double* sums, const b200_tensor* dpred,
                 float grad_scale, int32_t mode, void* stream);
/* softmax cross-entropy over channels (metrics.py:546, 581-586): target int64 class per voxel.  sums (double[3], zeroed by the
 * caller): [0] += loss over the counted voxels, [1] += counted voxels (label != ignore_index), [2] += voxels whose label is
 * neither a class nor ignore_index (torch asserts on those; they are skipped here and reported). */
int b200_softmax_ce(const b200_tensor* logits, const int64_t* target, double* sums, const b200_tensor* dlogits,
                    float grad_scale, int64_t ignore_index, void* stream);
/* head activations (biapy/engine/base_workflow.py:1367-1470): sigmoid per channel or softmax over [c0, c1) */
int b200_softmax_channels(const b200_tensor* x, const b200_tensor* y, int32_t c0, int32_t c1, void* stream);

/* ------------------------------------------------------------------------------------ batched weight re-packing
 * A training pass re-packs every fp32 master weight for the tensor-core kernels (b200_pack_conv_weight, _xfold, b200_pack_convT_weight:
 * 66 launches for the config-[1] network) and un-packs every weight gradient (b200_unpack_conv_wgrad, b200_unpack_convT_wgrad: 32);
 * b200_pack_batch runs any mix of those element mappings in ONE launch per 48 jobs.  `jobs` is a HOST array; src / dst are
 * device pointers.  kind: 0 plain pack, 1 x-folded pack, 2 transposed-conv pack (flip = for_dgrad), 3 un-pack of a conv weight (5: x-line pack, after 4)
 * gradient, 4 un-pack of a transposed-conv weight gradient (flip = accumulate into dst); (cout, cin, kd, kh, kw) as in the
 * single-job calls -- for kinds 2 / 4 `cout` carries the transposed conv's Cin and `cin` its Cout, the argument order of
 * b200_pack_convT_weight(w, packed, dtype, cin, cout, taps, ...).  block_begin / n_blocks / total are filled in by the call. */
typedef struct b200_pack_job {
  const float* src;
  void* dst;
  int32_t kind;
  int32_t cout, cin;
  int32_t kd, kh, kw;
  int32_t flip;
  int32_t block_begin, n_blocks;
  int64_t total;
} b200_pack_job;
int b200_pack_batch(const b200_pack_job* jobs, int32_t n_jobs, int32_t dtype, void* stream);

/* --------------------------------------------------------------------------------------------- optimiser
 * AdamW / SGD on a flat fp32 buffer (replaces timm create_optimizer_v2 -> torch.optim, engine/__init__.py:62-68). */
int b200_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int64_t step, float grad_scale, void* stream);
/* torch.optim.Adam (timm create_optimizer_v2('adam')): the weight decay is an L2 term added to the gradient. */
int b200_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int64_t step, float grad_scale, void* stream);
/* torch.optim.SGD incl. Nesterov momentum (timm's 'sgd' = SGD(momentum=0.9, nesterov=True), engine/__init__.py:58-70). */
int b200_sgd_step(float* p, const float* g, float* mom, int64_t n, float lr, float momentum, float weight_decay,
                  int32_t first_step, float grad_scale, int32_t nesterov, void* stream);
/* The same three optimisers with every hyper-parameter read from device memory, so that the update can sit inside a captured
 * CUDA graph while BiaPy's per-iteration LR schedule (train_engine.py:117-123) rewrites `hp` between replays.
 * kind: 0 AdamW, 1 Adam, 2 SGD.  hp (float[16]): lr, beta1, beta2, eps, weight_decay, momentum, nesterov, grad_scale, clip_norm.
 * gsq (nullable): sum of squares of g from b200_sumsq -- clip_grad_norm_ (train_engine.py:174-176) and the overflow test of
 * the fp16 engine (a non-finite gradient skips the update, as torch's GradScaler does).  denom (nullable): a divisor of the
 * gradient that only exists on the device (Noise2Void mask count, counted cross-entropy voxels).  state (int64[2]): steps taken,
 * steps skipped.  derived (float[8]): scratch written by the launch. */
int b200_optim_step_dev(int32_t kind, float* p, const float* g, float* m, float* v, int64_t n, const float* hp,
                        const double* gsq, const double* denom, int64_t* state, float* derived, void* stream);
/* dst[0..n) = values[0..n) (n <= 16) in stream order: the values travel as kernel arguments, so the host array may be reused
 * at once (how per-iteration hyper-parameters reach b200_optim_step_dev without a pinned staging buffer). */
/* dst[0 .. bytes) = 0 on the stream (cudaMemsetAsync: a memset node inside a captured graph): the gradient buffers and accumulators the
 * engine clears once per pass (reference: optimizer.zero_grad(), train_engine.py:106-203)                                          */
int b200_memset_zero(void* dst, int64_t bytes, void* stream);
int b200_write_floats(float* dst, const float* values, int32_t n, void* stream);
/* g *= mul / *denom with the divisor read on the device (per-rank mean of a masked loss before the gradient all-reduce). */
int b200_scale_by_dev(float* g, int64_t n, const double* denom, float mul, void* stream);
/* sum of squares of a flat fp32 buffer (for clip_grad_norm_, train_engine.py:174-176): out[0] (double) += ...   */
int b200_sumsq(const float* g, int64_t n, double* out, void* stream);

/* ------------------------------------------------------------------------------------------ self tests (GPU)
 * tcgen05 / TMA plumbing check: runs one small GEMM through every descriptor mode the conv kernels use and
 * compares against a SIMT reference on device.  Returns the number of mismatching modes (0 = all good).       */
int b200_umma_selftest(int32_t verbose, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BIAPY_B200_H */
