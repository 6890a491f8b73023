"""ctypes binding of the C ABI declared in ``include/biapy_b200.h``.

There is no CPU fallback: if the shared library is missing this module raises at import of the first symbol,
and every device entry point fails loudly when CUDA is unavailable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libbiapy_b200.so")

F32, BF16, F16, U8, U16 = 0, 1, 2, 3, 4
ACT = {None: 0, "none": 0, "linear": 0, "relu": 1, "elu": 2, "silu": 3, "leaky_relu": 4, "gelu": 5, "tanh": 6,
       "sigmoid": 7, "softplus": 8}
PAD_MODE = {"zeros": 0, "constant": 0, "reflect": 1, "symmetric": 2, "edge": 3, "wrap": 4}
IMPL_AUTO, IMPL_SIMT, IMPL_UMMA, IMPL_XFOLD = 0, 1, 2, 3


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("n", C.c_int32), ("d", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("c", C.c_int32), ("ld", C.c_int64)]


class AxisPlan(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("dim", "patch", "pad", "step", "n", "last", "core", "ov_px")]


class ChunkGrid(C.Structure):
    _fields_ = [("dim", C.c_int64 * 3), ("crop", C.c_int64 * 3), ("pad", C.c_int64 * 3), ("step", C.c_int64 * 3),
                ("vols", C.c_int64 * 3), ("z_vol_start", C.c_int64), ("z_vol_end", C.c_int64), ("total", C.c_int64)]


class B200Error(RuntimeError):
    pass


_lib = None
MISSING = []

_P = C.c_void_p
_I = C.c_int32
_L = C.c_int64
_F = C.c_float
_D = C.c_double
_T = C.POINTER(Tensor)

# name -> (restype, argtypes).  Must list every symbol of include/biapy_b200.h (tests/test_cabi.py checks).
SIGNATURES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_version": (_I, []),
    "b200_device_info": (_I, [_I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_L)]),
    "b200_plan_axis": (_I, [_L, _L, _L, C.c_double, C.POINTER(AxisPlan)]),
    "b200_axis_start": (_L, [C.POINTER(AxisPlan), _L, _I]),
    "b200_spline_window_1d": (_I, [_L, _L, C.POINTER(_F)]),
    "b200_crop_gather": (_I, [_P, _I, _L, _L, _L, _L, _P, _L, _L, _L, _P, _L, _P, _L, _P, _L, _L, _L, _L, _I, _P]),
    "b200_crop_gather_range": (_I, [_P, _I, _L, _L, _L, _L, _P, _L, _L, _L, _P, _L, _P, _L, _P, _L, _L, _L, _L, _I, _L, _L, _L, _L,
                                    _P]),
    "b200_overlap_add": (_I, [_P, _I, _P, _I, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _P, _L, _P, _L, _P, _L,
                              _P, _P, _P, _P]),
    "b200_overlap_add_slab": (_I, [_P, _I, _P, _I, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _P, _L, _P, _L, _P, _L,
                                   _P, _P, _P, _L, _L, _P]),
    "b200_chunk_grid_plan": (_I, [C.POINTER(_L), C.POINTER(_L), C.POINTER(_L), _L, _L, C.POINTER(ChunkGrid)]),
    "b200_chunk_patch_coords": (_I, [C.POINTER(ChunkGrid), _L, C.POINTER(_L)]),
    "b200_chunk_extract": (_I, [_P, _I, _L, _L, _L, _L, _P, _L, _L, _L, _L, _P, _P]),
    "b200_chunk_insert": (_I, [_P, _I, _L, _L, _L, _L, _L, _P, _I, _L, _L, _L, _P, _I, _P]),
    "b200_image_stats": (_I, [_P, _I, _L, _I, C.POINTER(_F), _P, _P]),
    "b200_edge_hist": (_I, [_P, _L, _P, _I, _P, _P]),
    "b200_select_hist": (_I, [_P, _I, _L, _I, _I, _I, _I, C.c_uint32, _I, _P, _P]),
    "b200_image_norm_apply": (_I, [_P, _I, _L, _I, C.POINTER(_F), _P, _P]),
    "b200_image_denorm_apply": (_I, [_P, _L, _I, C.POINTER(_D), _P, _I, _P]),
    "b200_binarize": (_I, [_P, _L, _F, _P, _P]),
    "b200_argmax_channels": (_I, [_P, _L, _I, _P, _I, _P]),
    "b200_orient_apply": (_I, [_T, _T, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), _I, _P]),
    "b200_orient_reduce": (_I, [_T, C.POINTER(_I), C.POINTER(_I), _I, C.POINTER(_I), _T, _P]),
    "b200_pack_conv_weight": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "b200_conv_fprop": (_I, [_T, _P, _P, _T, _T, _I, _I, _I, _I, _I, _P]),
    "b200_conv_wgrad": (_I, [_T, _T, _P, _P, _I, _I, _I, _I, _P]),
    "b200_conv_fprop_stats": (_I, [_T, _P, _P, _T, _I, _I, _I, _I, _P, C.POINTER(_I), _P]),
    "b200_pack_conv_weight_xfold": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "b200_conv_impl_query": (_I, [_T, _T, _I, _I, _I, _I]),
    "b200_conv_xline_supported": (_I, [_T, _T, _I, _I, _I]),
    "b200_pack_conv_weight_xline": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200_conv_fprop_xline": (_I, [_T, _P, _P, _T, _I, _P, _P, _I, _T, _P, _P]),
    "b200_xline_selftest": (_I, [C.POINTER(_D), _I, _P]),
    "b200_conv_wgrad_xline_supported": (_I, [_T, _T, _I, _I, _I]),
    "b200_conv_wgrad_xline": (_I, [_T, _T, _P, _P, _P]),
    "b200_unpack_conv_wgrad": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200_convT_fprop": (_I, [_T, _P, _P, _T, _I, _I, _I, _P]),
    "b200_convT_dgrad": (_I, [_T, _P, _T, _I, _I, _I, _I, _P]),
    "b200_convT_wgrad": (_I, [_T, _T, _P, _P, _I, _I, _I, _P]),
    "b200_convT_tc_supported": (_I, [_T, _T, _I, _I, _I]),
    "b200_pack_convT_weight": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "b200_convT_fprop_tc": (_I, [_T, _P, _P, _T, _I, _I, _I, _P]),
    "b200_convT_dgrad_tc": (_I, [_T, _P, _T, _I, _I, _I, _I, _P]),
    "b200_convT_wgrad_tc": (_I, [_T, _T, _P, _P, _I, _I, _I, _P]),
    "b200_unpack_convT_wgrad": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "b200_maxpool_fwd": (_I, [_T, _T, _I, _I, _I, _P]),
    "b200_maxpool_bwd": (_I, [_T, _T, _T, _T, _I, _I, _I, _I, _P]),
    "b200_maxpool_bwd_to_ok": (_I, [_T, _T, _T, _T, _I, _I, _I]),
    "b200_maxpool_bwd_to": (_I, [_T, _T, _T, _T, _I, _I, _I, _P]),
    "b200_channel_sums": (_I, [_T, _P, _P]),
    "b200_norm_finalize": (_I, [_P, _I, _I, _I, _L, _I, _P, _P, _F, _P, _P, _P, _P, _P]),
    "b200_scale_shift_act": (_I, [_T, _P, _P, _I, _T, _P]),
    "b200_norm_act_bwd_reduce": (_I, [_T, _T, _P, _P, _I, _P, _P, _I, _P, _P]),
    "b200_dropout": (_I, [_T, _T, _F, _P, _L, _I, _P]),
    "b200_upsample_linear_fwd": (_I, [_T, _T, _P]),
    "b200_upsample_linear_bwd": (_I, [_T, _T, _I, _P]),
    "b200_bn_update_running": (_I, [_P, _P, _F, _D, _F, _P, _P, _I, _P]),
    "b200_bn_eval_coeffs": (_I, [_P, _P, _P, _P, _F, _I, _I, _P, _P, _P]),
    "b200_norm_bwd_finalize": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _L, _I, _P, _P, _P, _P]),
    "b200_norm_bwd_finalize_sums": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _L, _I, _P, _P, _P, _P, _P, _P]),
    "b200_sums_through_pointwise": (_I, [_P, _P, _P, _I, _I, _P]),
    "b200_norm_act_bwd_apply": (_I, [_T, _T, _I, _P, _T, _I, _P]),
    "b200_norm_silu_fast_ok": (_I, [_T, _T, _T]),
    "b200_scale_shift_silu_fast": (_I, [_T, _P, _P, _T, _P]),
    "b200_norm_silu_bwd_reduce_g": (_I, [_T, _T, _P, _P, _I, _P, _P, _P, _I, _P]),
    "b200_norm_silu_bwd_apply_fast": (_I, [_T, _T, _P, _T, _I, _P]),
    "b200_norm_bwd_apply_g": (_I, [_T, _T, _P, _T, _I, _P]),
    "b200_act_bwd": (_I, [_T, _T, _I, _T, _I, _P]),
    "b200_binary": (_I, [_T, _T, _T, _I, _P]),
    "b200_gate_bwd": (_I, [_T, _T, _T, _T, _T, _I, _P]),
    "b200_relu_mask_bwd": (_I, [_T, _T, _T, _P]),
    "b200_convert": (_I, [_T, _T, _P]),
    "b200_bce_logits": (_I, [_T, _P, _P, _T, _F, _P]),
    "b200_n2v_mse": (_I, [_T, _P, _P, _T, _F, _I, _P]),
    "b200_softmax_ce": (_I, [_T, _P, _P, _T, _F, _L, _P]),
    "b200_softmax_channels": (_I, [_T, _T, _I, _I, _P]),
    "b200_adamw_step": (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _L, _F, _P]),
    "b200_adam_step": (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _L, _F, _P]),
    "b200_sgd_step": (_I, [_P, _P, _P, _L, _F, _F, _F, _I, _F, _I, _P]),
    "b200_optim_step_dev": (_I, [_I, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P]),
    "b200_sumsq": (_I, [_P, _L, _P, _P]),
    "b200_pack_batch": (_I, [_P, _I, _I, _P]),
    "b200_memset_zero": (_I, [_P, _L, _P]),
    "b200_write_floats": (_I, [_P, C.POINTER(_F), _I, _P]),
    "b200_scale_by_dev": (_I, [_P, _L, _P, _F, _P]),
    "b200_umma_selftest": (_I, [_I, _P]),
}


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} is missing: build it with `python -m biapy_b200.build` "
                "(biapy_b200 has no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(l, name)
            except AttributeError:
                MISSING.append(name)     # tests/test_cabi.py requires this list to be empty
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().b200_last_error().decode("utf-8", "replace")
        raise B200Error(f"{what or 'biapy_b200 call'} failed ({status}): {msg}")


def call(name: str, *args):
    """Call a status-returning entry point and raise B200Error on failure."""
    check(getattr(lib(), name)(*args), name)


# ------------------------------------------------------------------------------------------ torch helpers
def torch_dtype_code(dt) -> int:
    import torch
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float16:
        return F16
    if dt == torch.uint8:
        return U8
    if dt == torch.uint16:
        return U16
    raise B200Error(f"unsupported dtype {dt}")


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise B200Error(f"{what} must live on a CUDA device: biapy_b200 has no CPU path")


def as_tensor(t) -> Tensor:
    """View a channels-last torch tensor (N,D,H,W,C) (dense in N,D,H,W; C contiguous; arbitrary voxel pitch)."""
    assert t.dim() == 5, t.shape
    require_cuda(t)
    n, d, h, w, c = t.shape
    sn, sd, sh, sw, sc = t.stride()
    if c > 1 and sc != 1:
        raise B200Error(f"channel stride must be 1, got strides {t.stride()}")
    ld = sw if w > 1 else (sh // max(w, 1) if h > 1 else (sd // max(h * w, 1) if d > 1 else (sn // max(d * h * w, 1) if n > 1 else c)))
    # dense in the voxel index: stride(w)=ld, stride(h)=w*ld, stride(d)=h*w*ld, stride(n)=d*h*w*ld
    exp = (d * h * w * ld, h * w * ld, w * ld, ld)
    got = (sn, sd, sh, sw)
    for dim, e, g in zip((n, d, h, w), exp, got):
        if dim > 1 and e != g:
            raise B200Error(f"tensor is not voxel-dense: shape {tuple(t.shape)} strides {t.stride()}")
    return Tensor(t.data_ptr(), torch_dtype_code(t.dtype), n, d, h, w, c, ld)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
