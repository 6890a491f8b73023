"""Thin torch-tensor front end over the C ABI (``include/biapy_b200.h``).

Every function takes channels-last CUDA tensors ``(N, D, H, W, C)`` (a channel slice of a wider buffer is fine)
and launches hand-written CUDA on the current torch stream.  PyTorch is used for memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT, as_tensor, call, stream_ptr

# kernel launch counter (bench.py reports it as `gpu_launches`)
LAUNCHES = 0


# optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline):
# PROFILE = {} enables it; entries are name -> [(start_event, end_event, flops, bytes)]
PROFILE = None
PROFILE_SHAPES = False     # append the layer shape to the conv labels (bench.py --detail)


class ZeroArena:
    """Zero-initialised scratch for one training pass.  The accumulators of a step (channel sums, backward sums, packed
    weight-gradient buffers: ~70 small tensors) each cost a fill kernel when allocated with ``torch.zeros``; between
    :meth:`begin` and :meth:`end` they are carved out of ONE buffer that is cleared by a single fill.  The buffer grows to the
    high-water mark of the previous pass (the first pass falls back to ``torch.zeros``)."""

    def __init__(self):
        self.buf = None
        self._retired = []          # outgrown buffers stay alive: a captured CUDA graph may still point into them
        self.off = self.need = 0

    def begin(self, device):
        """Make this the active arena (module-level `ARENA`) and clear it."""
        global ARENA
        size = (self.need + 4095) // 4096 * 4096
        if size and (self.buf is None or self.buf.numel() < size or self.buf.device != torch.device(device)):
            if self.buf is not None:
                self._retired.append(self.buf)
            self.buf = torch.empty(size, dtype=torch.uint8, device=device)
        if self.buf is not None:
            zero_(self.buf)
        self.off = self.need = 0
        ARENA = self

    def end(self):
        global ARENA
        ARENA = None

    def take(self, numel: int, dtype: torch.dtype, device) -> Optional[torch.Tensor]:
        nbytes = (numel * dtype.itemsize + 255) // 256 * 256
        self.need += nbytes
        if self.buf is None or self.off + nbytes > self.buf.numel() or self.buf.device != torch.device(device):
            return None
        t = self.buf[self.off:self.off + nbytes].view(dtype)[:numel]
        self.off += nbytes
        return t


ARENA: Optional[ZeroArena] = None      # the arena of the training pass in flight (set by ZeroArena.begin / end)


def zero_(t: torch.Tensor) -> torch.Tensor:
    """t[...] = 0 for a CUDA tensor whose elements are one contiguous block (cudaMemsetAsync on the current stream)."""
    if not t.is_cuda or not t.is_contiguous():
        return t.zero_()
    call("b200_memset_zero", _ptr(t), t.numel() * t.element_size(), stream_ptr())      # a memset, not a kernel: not counted in LAUNCHES
    return t


def zeros(numel: int, dtype: torch.dtype, device) -> torch.Tensor:
    """Zero-filled 1-D scratch tensor: a slice of the step's arena when one is active, else ``torch.zeros``."""
    if ARENA is not None:
        t = ARENA.take(numel, dtype, device)
        if t is not None:
            return t
    return torch.zeros(numel, dtype=dtype, device=device)


def _timed_call(label, flops, nbytes, name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, *args)
    e1.record()
    PROFILE.setdefault(label, []).append((e0, e1, flops, nbytes))


def _launch(name, *args, shape=None):
    global LAUNCHES
    LAUNCHES += 1
    if PROFILE is None:
        return call(name, *args)
    label = name[5:]
    if PROFILE_SHAPES and shape is not None:
        label += f" c{shape[-1]} @{shape[1]}x{shape[2]}x{shape[3]}"
    _timed_call(label, 0.0, 0.0, name, *args)


def _launch_timed(label, flops, nbytes, name, *args):
    global LAUNCHES
    LAUNCHES += 1
    if PROFILE is None:
        return call(name, *args)
    _timed_call(label, flops, nbytes, name, *args)


# ------------------------------------------------------------------------------------------ batched re-packing
class PackJob(C.Structure):
    """Mirror of ``b200_pack_job`` (include/biapy_b200.h)."""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("kind", C.c_int32), ("cout", C.c_int32), ("cin", C.c_int32),
                ("kd", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("flip", C.c_int32), ("block_begin", C.c_int32),
                ("n_blocks", C.c_int32), ("total", C.c_int64)]


PACK_PLAIN, PACK_XFOLD, PACK_CONVT, UNPACK_WGRAD, UNPACK_CONVT_WGRAD, PACK_XLINE = 0, 1, 2, 3, 4, 5
# set by the single-job pack functions: (kind, src tensor, dst tensor, cout, cin, kd, kh, kw, flip) of the launch just made --
# the Tape copies it into the Trainer's pack plan so that later passes replay all packs as one launch
LAST_PACK = None
# list while a Trainer pass is in flight: weight-gradient un-packs are queued here and flushed as one launch after backward
UNPACK_QUEUE = None


def pack_batch(jobs, dtype: torch.dtype):
    """jobs: [(kind, src, dst, cout, cin, kd, kh, kw, flip)] with CUDA tensors src / dst; one launch per 48 jobs."""
    if not jobs:
        return
    arr = (PackJob * len(jobs))()
    for a, (kind, src, dst, cout, cin, kd, kh, kw, flip) in zip(arr, jobs):
        a.src, a.dst, a.kind, a.cout, a.cin, a.kd, a.kh, a.kw, a.flip = src.data_ptr(), dst.data_ptr(), kind, cout, cin, kd, kh, kw, flip
    global LAUNCHES
    LAUNCHES += (len(jobs) - 1) // 48
    _launch("b200_pack_batch", arr, len(jobs), _lib.torch_dtype_code(dtype), stream_ptr())


def flush_unpacks():
    """Launch the queued weight-gradient un-packs (see UNPACK_QUEUE) and empty the queue."""
    global UNPACK_QUEUE
    q = UNPACK_QUEUE
    if q:
        UNPACK_QUEUE = []
        pack_batch([j[:9] for j in q], torch.bfloat16)      # fp32 -> fp32 jobs: the dtype only selects the (unused) pack type


def queue_float_add(src: torch.Tensor, dst: torch.Tensor):
    """dst += src (contiguous fp32 vectors, e.g. a bias gradient shared by two layers): rides in the batched un-pack launch of a
    Trainer pass, a launch of its own otherwise."""
    n = src.numel()
    assert dst.numel() == n and src.dtype == dst.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
    if UNPACK_QUEUE is not None:
        UNPACK_QUEUE.append((UNPACK_WGRAD, src, dst, n, 1, 1, 1, 1, 1))
    else:
        _launch("b200_unpack_conv_wgrad", _ptr(src), _ptr(dst), n, 1, 1, 1, stream_ptr())


def sums_through_pointwise(w: torch.Tensor, dysum: torch.Tensor, xsum: torch.Tensor):
    """xsum (Cin,) += W^T dysum for the pointwise convolution weight W (Cout, Cin, 1, 1, 1) fp32."""
    cout, cin = w.shape[0], w.shape[1]
    assert w.is_contiguous() and w.dtype == torch.float32 and dysum.numel() == cout and xsum.numel() == cin
    _launch("b200_sums_through_pointwise", _ptr(w), _ptr(dysum), _ptr(xsum), cout, cin, stream_ptr())


def conv_impl_query(x, y, k, wgrad=False) -> int:
    """Best kernel family for these operands: IMPL_XFOLD / IMPL_UMMA / IMPL_SIMT."""
    return _lib.lib().b200_conv_impl_query(_ref(x), _ref(y), k[0], k[1], k[2], 1 if wgrad else 0)


def _family(kind: str, x, y, k, wgrad, impl) -> str:
    """Profile label of a convolution launch.  Families follow the roofline that bounds them: 3x3(x3) layers with
    >= 16 input channels are tensor-bound `conv_<kind>_<impl>`; pointwise layers (k = 1: 12 FLOP per byte at 16 <-> 48 channels)
    and image-fed layers (Cin < 16) are HBM streams and get their own families so that they do not dilute the tensor figure."""
    name = _impl_name(x, y, k, wgrad, impl)
    if tuple(k) == (1, 1, 1):
        return f"conv1x1_{kind}_{name}"
    if x.shape[-1] < 16:
        return f"conv_image_{kind}_{name}"
    return f"conv_{kind}_{name}"


def _impl_name(x, y, k, wgrad, impl):
    if impl == _lib.IMPL_AUTO:
        impl = conv_impl_query(x, y, k, wgrad)
        if impl == _lib.IMPL_XFOLD:
            impl = _lib.IMPL_UMMA        # AUTO never selects the x-folded kernel (different weight packing)
    return {_lib.IMPL_UMMA: "umma", _lib.IMPL_XFOLD: "xfold"}.get(impl, "simt")


def _ref(t):
    return C.byref(as_tensor(t)) if t is not None else None


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else None


def empty_like_cl(t: torch.Tensor, channels: Optional[int] = None, dtype=None) -> torch.Tensor:
    shp = list(t.shape)
    if channels is not None:
        shp[-1] = channels
    return torch.empty(shp, dtype=dtype or t.dtype, device=t.device)


# ---------------------------------------------------------------------------------------------- convolution
def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype, flip_transpose: bool) -> torch.Tensor:
    """w: (Cout, Cin, *k) fp32 parameter -> packed [Cout][tap][Cin] (or [Cin][flipped tap][Cout])."""
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    cout, cin = w.shape[:2]
    k = tuple(w.shape[2:])
    if len(k) == 2:
        k = (1,) + k
    out = torch.empty(w.numel(), dtype=dtype, device=w.device)
    _launch("b200_pack_conv_weight", _ptr(w), _ptr(out), _lib.torch_dtype_code(dtype), cout, cin, k[0], k[1], k[2],
            1 if flip_transpose else 0, stream_ptr())
    global LAST_PACK
    LAST_PACK = (PACK_PLAIN, w, out, cout, cin, k[0], k[1], k[2], 1 if flip_transpose else 0)
    return out


def xfold_window(cin: int, kw: int) -> Tuple[int, int]:
    """(xoff, kxp) of the x-folded row window -- mirrors `xfold_geom` in csrc/conv_umma.cu."""
    pw = kw // 2
    if cin % 16 == 0:
        return 0, (3 + kw) * cin
    if cin not in (2, 4, 8):
        raise _lib.B200Error(f"x-folded kernels take Cin in (2, 4, 8) or a multiple of 16, got {cin}")
    xo = 0
    while ((pw + xo) * cin * 2) % 16:
        xo += 1
    return xo, ((xo + 4 + 2 * pw) * cin + 31) // 32 * 32


def pack_conv_weight_xfold(w: torch.Tensor, dtype: torch.dtype, flip_transpose: bool) -> torch.Tensor:
    """w: (Cout, Cin, kd, kh, 3) or (Cout, Cin, kh, 3) fp32 -> block-Toeplitz packing of the x-folded kernel."""
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    cout, cin = w.shape[:2]
    k = tuple(w.shape[2:])
    if len(k) == 2:
        k = (1,) + k
    assert k[2] in (1, 3)
    co_l, ci_l = (cin, cout) if flip_transpose else (cout, cin)
    out = torch.empty(4 * co_l * k[0] * k[1] * xfold_window(ci_l, k[2])[1], dtype=dtype, device=w.device)
    _launch("b200_pack_conv_weight_xfold", _ptr(w), _ptr(out), _lib.torch_dtype_code(dtype), cout, cin, k[0], k[1], k[2],
            1 if flip_transpose else 0, stream_ptr())
    global LAST_PACK
    LAST_PACK = (PACK_XFOLD, w, out, cout, cin, k[0], k[1], k[2], 1 if flip_transpose else 0)
    return out


def conv_xline_supported(x, y, k) -> bool:
    """True when the x-line kernel (csrc/conv_xline.cu) takes these operands: 3x3x3, W = 128, (Cin, Cout) in (16 | 48, 16) or
    (16, 48) -- the latter without accumulation, statistics or fusion -- dense 16-bit channels-last input, and B200_XLINE != 0."""
    if x.dtype == torch.float32 or tuple(k) != (3, 3, 3):
        return False
    return bool(_lib.lib().b200_conv_xline_supported(_ref(x), _ref(y), 3, 3, 3))


def pack_conv_weight_xline(w: torch.Tensor, dtype: torch.dtype, flip_transpose: bool) -> torch.Tensor:
    """w: (Cout, Cin, 3, 3, 3) fp32 -> rotated (48 x 16) tiles of the x-line kernel (`b200_pack_conv_weight_xline`)."""
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    cout, cin = w.shape[:2]
    co_l, ci_l = (cin, cout) if flip_transpose else (cout, cin)
    out = torch.empty(27 * (ci_l // 16) * co_l * 48, dtype=dtype, device=w.device)
    _launch("b200_pack_conv_weight_xline", _ptr(w), _ptr(out), _lib.torch_dtype_code(dtype), cout, cin, 1 if flip_transpose else 0,
            stream_ptr())
    global LAST_PACK
    LAST_PACK = (PACK_XLINE, w, out, cout, cin, 3, 3, 3, 1 if flip_transpose else 0)
    return out


def conv_fprop_xline(x, w_packed_xline, bias, y, accumulate=False, scale=None, shift=None, fuse: int = 0, a_out=None, sums=None):
    """x-line convolution; fuse = 1 / 2: silu(x * scale + shift) (exact / one-MUFU chain) applied to the input on the operand
    path, `a_out` receives the activated input (`b200_conv_fprop_xline`)."""
    label = flops = nbytes = None
    if PROFILE is not None:
        vox = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]
        flops = 2.0 * vox * x.shape[-1] * y.shape[-1] * 27
        nbytes = vox * (x.shape[-1] * (2 if a_out is not None else 1) + y.shape[-1] * (2 if accumulate else 1)) * x.element_size()
        label = "conv_fprop_xline" + ("_gn_silu" if fuse else "")
        if PROFILE_SHAPES:
            label += (f" {x.shape[-1]}->{y.shape[-1]} k333 @{x.shape[1]}x{x.shape[2]}x{x.shape[3]}"
                      + (" +stats" if sums is not None else ""))
    _launch_timed(label, flops, nbytes, "b200_conv_fprop_xline", _ref(x), _ptr(w_packed_xline), _ptr(bias), _ref(y),
                  1 if accumulate else 0, _ptr(scale), _ptr(shift), int(fuse), _ref(a_out), _ptr(sums), stream_ptr())
    return y


def xline_selftest(verbose: int = 0) -> float:
    err = C.c_double(0.0)
    call("b200_xline_selftest", C.byref(err), int(verbose), stream_ptr())
    return float(err.value)


def conv_fprop(x, w_packed, bias, y, k: Sequence[int], residual=None, accumulate=False, impl=_lib.IMPL_AUTO):
    label = flops = nbytes = None
    if PROFILE is not None:
        vox = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]
        flops = 2.0 * vox * x.shape[-1] * y.shape[-1] * k[0] * k[1] * k[2]
        nbytes = vox * (x.shape[-1] + y.shape[-1] * (2 if accumulate else 1)) * x.element_size()
        label = _family("fprop", x, y, k, False, impl)
        if PROFILE_SHAPES:
            label += f" {x.shape[-1]}->{y.shape[-1]} k{k[0]}{k[1]}{k[2]} @{x.shape[1]}x{x.shape[2]}x{x.shape[3]}"
    _launch_timed(label, flops, nbytes, "b200_conv_fprop", _ref(x), _ptr(w_packed), _ptr(bias), _ref(residual), _ref(y), k[0], k[1], k[2],
            1 if accumulate else 0, impl, stream_ptr())
    return y


def conv_fprop_stats(x, w_packed_xfold, bias, y, k: Sequence[int], sums: torch.Tensor, accumulate=False) -> bool:
    """x-folded convolution with the channel statistics of its output fused into the epilogue (`b200_conv_fprop_stats`):
    ``sums[n][c] += (sum y, sum y^2)`` -- what `b200_channel_sums` would compute from y.  Returns True when the kernel
    produced them; on False `sums` is untouched (the convolution still ran)."""
    label = flops = nbytes = None
    if PROFILE is not None:
        vox = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]
        flops = 2.0 * vox * x.shape[-1] * y.shape[-1] * k[0] * k[1] * k[2]
        nbytes = vox * (x.shape[-1] + y.shape[-1] * (2 if accumulate else 1)) * x.element_size()
        label = _family("fprop", x, y, k, False, _lib.IMPL_XFOLD)
        if PROFILE_SHAPES:
            label += f" {x.shape[-1]}->{y.shape[-1]} k{k[0]}{k[1]}{k[2]} @{x.shape[1]}x{x.shape[2]}x{x.shape[3]} +stats"
    applied = C.c_int32(0)
    _launch_timed(label, flops, nbytes, "b200_conv_fprop_stats", _ref(x), _ptr(w_packed_xfold), _ptr(bias), _ref(y), k[0], k[1], k[2],
                  1 if accumulate else 0, _ptr(sums), C.byref(applied), stream_ptr())
    return bool(applied.value)


# B200_XLINE_WGRAD = 0 keeps the 16-output-channel 3x3x3 weight gradients at W = 128 on the x-folded kernels
XLINE_WGRAD = os.environ.get("B200_XLINE_WGRAD", "1") != "0"
# input-channel counts that take it (128^3 x 4, fp16): 16 -> 16 0.22 ms against 0.39 x-folded; 48 -> 16 (three launches, one per
# 16-channel group, each transposing dY again) 0.65 against 0.85
XLINE_WGRAD_CIN = tuple(int(v) for v in os.environ.get("B200_XLINE_WGRAD_CIN", "16,48").split(",") if v)

_WGRAD_N = (256, 128, 64, 32, 16)        # output-channel widths of the tcgen05 weight-gradient kernels (one N tile each)


def _wgrad_cuts(cout: int):
    """Output-channel slices [(a, b), ...] that the tensor-core weight-gradient kernels take: Cout itself when it is one of their
    N tiles, else a greedy cover by them (512 -> 256 + 256, 384 -> 256 + 128, 48 -> 32 + 16)."""
    if cout in _WGRAD_N or cout % 16:
        return [(0, cout)]
    cuts, a = [], 0
    while a < cout:
        w = next(n for n in _WGRAD_N if n <= cout - a)
        cuts.append((a, a + w))
        a += w
    return cuts


def conv_wgrad(x, dy, cout: int, cin: int, k: Sequence[int], dw_out: torch.Tensor, dbias_out: Optional[torch.Tensor],
               accumulate=False, impl=_lib.IMPL_AUTO, defer_unpack: bool = True):
    """dw_out: (Cout, Cin, *k) fp32; dbias_out: (Cout,) fp32, must be zero-initialised unless accumulate.  Inside a Trainer pass
    the un-pack of the gradient into `dw_out` is queued (`UNPACK_QUEUE`) unless `defer_unpack` is False."""
    taps = k[0] * k[1] * k[2]
    packed = zeros(cout * taps * cin, torch.float32, x.device)
    if (impl == _lib.IMPL_AUTO and XLINE_WGRAD and x.dtype != torch.float32 and cout == 16 and tuple(k) == (3, 3, 3)
            and x.shape[3] == 128 and cin in XLINE_WGRAD_CIN
            and _lib.lib().b200_conv_wgrad_xline_supported(_ref(x), _ref(dy), 3, 3, 3)):
        # x-line weight gradient (csrc/conv_xline.cu): the contraction runs over the voxels of a line, the transposed activation
        # line sits in tensor memory, the 27-tap block is accumulated in tensor memory for the whole launch
        label = flops = nbytes = None
        if PROFILE is not None:
            vox = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]
            flops = 2.0 * vox * cin * cout * taps
            nbytes = vox * (cin + cout) * x.element_size()
            label = "conv_wgrad_xline"
            if PROFILE_SHAPES:
                label += f" {cin}->{cout} k333 @{x.shape[1]}x{x.shape[2]}x{x.shape[3]}"
        _launch_timed(label, flops, nbytes, "b200_conv_wgrad_xline", _ref(x), _ref(dy), _ptr(packed), _ptr(dbias_out), stream_ptr())
        global LAUNCHES
        LAUNCHES += cin // 16 - 1                     # the bias gradient rides in the first launch (a row of ones in the operand)
        if defer_unpack and UNPACK_QUEUE is not None and dw_out.is_contiguous():
            UNPACK_QUEUE.append((UNPACK_WGRAD, packed, dw_out, cout, cin, taps, 1, 1, 1 if accumulate else 0))
        else:
            _launch("b200_unpack_conv_wgrad", _ptr(packed), _ptr(dw_out), cout, cin, taps, 1 if accumulate else 0, stream_ptr())
        return
    cuts = [(0, cout)]
    if impl == _lib.IMPL_AUTO and x.dtype != torch.float32 and conv_impl_query(x, dy, k, True) == _lib.IMPL_SIMT:
        # Cout outside the kernels' N tiles (e.g. the 512-channel layers of BASELINE config[4]): one launch per slice of dy, each
        # writing its own rows of the packed [Cout][tap][Cin] gradient -- instead of the CUDA-core fallback for the whole layer
        c2 = _wgrad_cuts(cout)
        if len(c2) > 1 and all(conv_impl_query(x, dy[..., a:b], k, True) != _lib.IMPL_SIMT for a, b in c2):
            cuts = c2
    for a, b in cuts:
        dys = dy if len(cuts) == 1 else dy[..., a:b]
        label = flops = nbytes = None
        if PROFILE is not None:
            vox = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]
            flops = 2.0 * vox * cin * (b - a) * taps
            nbytes = vox * (cin + (b - a)) * x.element_size()
            label = _family("wgrad", x, dys, k, True, impl)
            if PROFILE_SHAPES:
                label += f" {cin}->{b - a} k{k[0]}{k[1]}{k[2]} @{x.shape[1]}x{x.shape[2]}x{x.shape[3]}"
        _launch_timed(label, flops, nbytes, "b200_conv_wgrad", _ref(x), _ref(dys), _ptr(packed[a * taps * cin:b * taps * cin]),
                      _ptr(None if dbias_out is None else dbias_out[a:b]), k[0], k[1], k[2], impl, stream_ptr())
    if defer_unpack and UNPACK_QUEUE is not None and dw_out.is_contiguous():
        UNPACK_QUEUE.append((UNPACK_WGRAD, packed, dw_out, cout, cin, taps, 1, 1, 1 if accumulate else 0))
    else:
        _launch("b200_unpack_conv_wgrad", _ptr(packed), _ptr(dw_out), cout, cin, taps, 1 if accumulate else 0, stream_ptr())


def convT_fprop(x, w, bias, y, s: Sequence[int]):
    _launch("b200_convT_fprop", _ref(x), _ptr(w), _ptr(bias), _ref(y), s[0], s[1], s[2], stream_ptr())
    return y


def convT_dgrad(dy, w, dx, s: Sequence[int], accumulate=False):
    _launch("b200_convT_dgrad", _ref(dy), _ptr(w), _ref(dx), s[0], s[1], s[2], 1 if accumulate else 0, stream_ptr())


def convT_wgrad(x, dy, dw, dbias, s: Sequence[int]):
    _launch("b200_convT_wgrad", _ref(x), _ref(dy), _ptr(dw), _ptr(dbias), s[0], s[1], s[2], stream_ptr())


def convT_tc_supported(x, y, s: Sequence[int]) -> bool:
    return bool(_lib.lib().b200_convT_tc_supported(_ref(x), _ref(y), s[0], s[1], s[2]))


def pack_convT_weight(w: torch.Tensor, dtype: torch.dtype, for_dgrad: bool) -> torch.Tensor:
    """w: (Cin, Cout, *s) fp32 -> [tap][Cout][Cin] (fprop) or [tap][Cin][Cout] (dgrad) in the engine dtype."""
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    cin, cout = w.shape[:2]
    taps = w[0, 0].numel()
    out = torch.empty(w.numel(), dtype=dtype, device=w.device)
    _launch("b200_pack_convT_weight", _ptr(w), _ptr(out), _lib.torch_dtype_code(dtype), cin, cout, taps, 1 if for_dgrad else 0,
            stream_ptr())
    global LAST_PACK
    LAST_PACK = (PACK_CONVT, w, out, cin, cout, taps, 1, 1, 1 if for_dgrad else 0)
    return out


def _convT_work(x, cout, s):
    vox = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]
    return 2.0 * vox * x.shape[-1] * cout * s[0] * s[1] * s[2]


def convT_fprop_tc(x, w_packed, bias, y, s: Sequence[int]):
    _launch_timed("convT_fprop_tc", _convT_work(x, y.shape[-1], s) if PROFILE is not None else 0, 0,
                  "b200_convT_fprop_tc", _ref(x), _ptr(w_packed), _ptr(bias), _ref(y), s[0], s[1], s[2], stream_ptr())
    return y


def convT_dgrad_tc(dy, w_packed_t, dx, s: Sequence[int], accumulate=False):
    _launch_timed("convT_dgrad_tc", _convT_work(dx, dy.shape[-1], s) if PROFILE is not None else 0, 0,
                  "b200_convT_dgrad_tc", _ref(dy), _ptr(w_packed_t), _ref(dx), s[0], s[1], s[2], 1 if accumulate else 0, stream_ptr())


def convT_wgrad_tc(x, dy, dw: torch.Tensor, dbias, s: Sequence[int], accumulate=False):
    """dw: (Cin, Cout, *s) fp32 parameter-layout gradient."""
    cin, cout = x.shape[-1], dy.shape[-1]
    taps = s[0] * s[1] * s[2]
    packed = zeros(taps * cout * cin, torch.float32, x.device)
    _launch_timed("convT_wgrad_tc", _convT_work(x, cout, s) if PROFILE is not None else 0, 0,
                  "b200_convT_wgrad_tc", _ref(x), _ref(dy), _ptr(packed), _ptr(dbias), s[0], s[1], s[2], stream_ptr())
    if UNPACK_QUEUE is not None and dw.is_contiguous():
        UNPACK_QUEUE.append((UNPACK_CONVT_WGRAD, packed, dw, cin, cout, taps, 1, 1, 1 if accumulate else 0))
    else:
        _launch("b200_unpack_convT_wgrad", _ptr(packed), _ptr(dw), cin, cout, taps, 1 if accumulate else 0, stream_ptr())


# ------------------------------------------------------------------------------------------------- dropout
def dropout(x, y, p: float, seed_dev: torch.Tensor, layer_id: int, accumulate=False):
    """y (+)= mask * x / (1 - p); the mask is a function of (seed_dev[0], layer_id, element index)."""
    _launch("b200_dropout", _ref(x), _ref(y), float(p), _ptr(seed_dev), int(layer_id), 1 if accumulate else 0, stream_ptr())
    return y


# ------------------------------------------------------------------------------------- linear up-sampling
def upsample_linear_fwd(x, y):
    """nn.Upsample(bilinear/trilinear, align_corners=False); integer scale factors = y dims / x dims."""
    _launch("b200_upsample_linear_fwd", _ref(x), _ref(y), stream_ptr())
    return y


def upsample_linear_bwd(dy, dx, accumulate=False):
    _launch("b200_upsample_linear_bwd", _ref(dy), _ref(dx), 1 if accumulate else 0, stream_ptr())


# ------------------------------------------------------------------------------------------------- pooling
def maxpool_fwd(x, y, p: Sequence[int]):
    _launch("b200_maxpool_fwd", _ref(x), _ref(y), p[0], p[1], p[2], stream_ptr())
    return y


def maxpool_bwd(x, y, dy, dx, p: Sequence[int], accumulate=False):
    _launch("b200_maxpool_bwd", _ref(x), _ref(y), _ref(dy), _ref(dx), p[0], p[1], p[2], 1 if accumulate else 0, stream_ptr())


def maxpool_bwd_to(x, dy, dx_in, dx_out, p: Sequence[int]) -> bool:
    """dx_out (dense) = (dx_in or 0) + routed dy; returns False (nothing launched) when the operands do not qualify."""
    if not _lib.lib().b200_maxpool_bwd_to_ok(_ref(x), _ref(dy), _ref(dx_in), _ref(dx_out), p[0], p[1], p[2]):
        return False
    _launch("b200_maxpool_bwd_to", _ref(x), _ref(dy), _ref(dx_in), _ref(dx_out), p[0], p[1], p[2], stream_ptr())
    return True


# ------------------------------------------------------------------------------------- normalisation / act
class NormStats:
    __slots__ = ("mean", "rstd", "scale", "shift", "groups", "batch_stats", "world", "sync_group", "sums")


def _sync_world(sync_group) -> int:
    """World size for synchronised batch statistics (1 = no collective)."""
    if sync_group is False or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return 1
    return torch.distributed.get_world_size(None if sync_group is True else sync_group)


def norm_stats(x, groups: int, gamma, beta, eps: float = 1e-5, batch_stats: bool = False, sync_group=False,
               sums: Optional[torch.Tensor] = None) -> NormStats:
    """`sync_group`: False, True (default process group) or a process group -- SyncBatchNorm: the per-rank channel sums
    are all-reduced before the statistics are finalised (equal per-rank batch sizes assumed, as under DDP).
    `sums`: the (N, C, 2) float64 channel sums of x when the producing convolution already reduced them in its epilogue."""
    n, d, h, w, c = x.shape
    if sums is None:
        sums = zeros(n * c * 2, torch.float64, x.device)
        _launch("b200_channel_sums", _ref(x), _ptr(sums), stream_ptr(), shape=x.shape)
    world = _sync_world(sync_group) if batch_stats else 1
    if world > 1:
        # batch statistics only need the sum over samples: reduce the (C, 2) totals, not the per-sample table
        tot = sums.view(n, c * 2).sum(0)
        torch.distributed.all_reduce(tot, group=None if sync_group is True else sync_group)
        sums = torch.zeros_like(sums)
        sums.view(n, c * 2)[0] = tot
    st = NormStats()
    st.groups, st.batch_stats = groups, batch_stats
    st.world, st.sync_group = world, sync_group
    st.sums = sums if world == 1 else None        # per-sample channel sums of x: the backward derives sum(dx) from them (norm_act_bwd)
    buf = torch.empty(2 * n * groups + 2 * n * c, dtype=torch.float32, device=x.device)
    st.mean, st.rstd = buf[: n * groups], buf[n * groups: 2 * n * groups]
    st.scale, st.shift = buf[2 * n * groups: 2 * n * groups + n * c], buf[2 * n * groups + n * c:]
    _launch("b200_norm_finalize", _ptr(sums), n, c, groups, d * h * w * world, 1 if batch_stats else 0, _ptr(gamma), _ptr(beta),
            eps, _ptr(st.mean), _ptr(st.rstd), _ptr(st.scale), _ptr(st.shift), stream_ptr())
    return st


def bn_update_running(st: NormStats, c: int, count: float, eps: float, momentum: float, running_mean, running_var):
    """running statistics of nn.BatchNorm from the batch statistics in `st` (unbiased variance, PyTorch semantics)."""
    _launch("b200_bn_update_running", _ptr(st.mean), _ptr(st.rstd), float(eps), float(count), float(momentum), _ptr(running_mean),
            _ptr(running_var), c, stream_ptr())


def bn_eval_stats(x, running_mean, running_var, gamma, beta, eps: float) -> NormStats:
    """eval-mode BatchNorm: scale/shift from the running statistics (no reduction over x)."""
    n, c = x.shape[0], x.shape[-1]
    st = NormStats()
    st.groups, st.batch_stats, st.world, st.sync_group = c, True, 1, False
    st.sums = None
    buf = torch.empty(2 * n * c, dtype=torch.float32, device=x.device)
    st.mean = st.rstd = None
    st.scale, st.shift = buf[: n * c], buf[n * c:]
    _launch("b200_bn_eval_coeffs", _ptr(running_mean), _ptr(running_var), _ptr(gamma), _ptr(beta), float(eps), n, c,
            _ptr(st.scale), _ptr(st.shift), stream_ptr())
    return st


# B200_NORM_FAST: 'auto' (default) = the one-MUFU SiLU chain for bf16 tensors and for the BACKWARD of fp16 tensors, '1' = fp16
# forward as well, '0' = off, 'bf16' = bf16 only.  tanh.approx has a relative error of 2^-11: below bf16's rounding step, comparable
# to fp16's -- the fp16 engine is the 1e-3 parity path for OUTPUTS and keeps the exact exponential in the forward pass; its
# gradients carry the 16-bit storage error of the whole backward chain (1e-2), two orders above what the approximate
# sigmoid of the derivative adds.
NORM_FAST = os.environ.get("B200_NORM_FAST", "auto").lower()
# backward form of the fast chain: 'g' = pass 1 leaves g = dy * act' in dy's place (needs dy_dead), 'recompute' = pass 2 evaluates
# the derivative again with the one-MUFU sigmoid (no extra write in pass 1)
# measured (config[1] step, profiles/README.md round 2): 'recompute' 14.79 ms, 'g' 14.83 ms -- the extra write of pass 1 costs what the
# second derivative saves; 'recompute' is the default because it leaves dy alone
NORM_BWD = os.environ.get("B200_NORM_BWD", "recompute").lower()


def norm_fast_ok(x, dy=None, dx=None, backward: bool = False) -> bool:
    if NORM_FAST in ("0", "off") or x.dtype == torch.float32:
        return False
    if x.dtype == torch.float16 and not (NORM_FAST == "1" or (backward and NORM_FAST == "auto")):
        return False
    return bool(_lib.lib().b200_norm_silu_fast_ok(_ref(x), _ref(dy), _ref(dx)))


def scale_shift_act(x, scale, shift, act: str, y):
    if act == "silu" and scale is not None and x.dtype == y.dtype and norm_fast_ok(x, y):
        _launch("b200_scale_shift_silu_fast", _ref(x), _ptr(scale), _ptr(shift), _ref(y), stream_ptr(), shape=x.shape)
        return y
    _launch("b200_scale_shift_act", _ref(x), _ptr(scale), _ptr(shift), ACT[act], _ref(y), stream_ptr(), shape=x.shape)
    return y


def norm_act_bwd(x, dy, st: NormStats, gamma, beta, act: str, dx, dgamma, dbeta, accumulate=False, dy_dead: bool = False,
                 dx_sums: Optional[torch.Tensor] = None):
    """`dy_dead`: nothing reads dy after this call (true for the tape's activation gradients) -- allows the fast SiLU chain,
    whose reduce pass leaves g = dy * act'(z) in dy's place for the apply pass.  `dx_sums`: (C,) fp32, += the sum over samples and
    voxels of the dx this call produces (its own contribution when accumulating), derived from the reductions -- needs `st.sums`."""
    n, d, h, w, c = x.shape
    red = zeros(n * c * 2, torch.float64, x.device)
    write_g = NORM_BWD != "recompute"
    fast = (dy_dead or not write_g) and act == "silu" and norm_fast_ok(x, dy, dx, backward=True)
    if fast:
        _launch("b200_norm_silu_bwd_reduce_g", _ref(x), _ref(dy), _ptr(st.mean), _ptr(st.rstd), st.groups, _ptr(gamma), _ptr(beta),
                _ptr(red), 1 if write_g else 0, stream_ptr(), shape=x.shape)
    else:
        _launch("b200_norm_act_bwd_reduce", _ref(x), _ref(dy), _ptr(st.mean), _ptr(st.rstd), st.groups, _ptr(gamma), _ptr(beta),
                ACT[act], _ptr(red), stream_ptr(), shape=x.shape)
    coef = torch.empty(n * c * 4, dtype=torch.float32, device=x.device)
    world = getattr(st, "world", 1) or 1
    if world > 1:
        # SyncBatchNorm: dgamma/dbeta are this rank's own sums (DDP averages them later); the dx coefficients need the
        # sums over all ranks (torch.nn.SyncBatchNorm backward does the same all-reduce of sum_dy, sum_dy_xmu)
        _launch("b200_norm_bwd_finalize", _ptr(red), _ptr(st.mean), _ptr(st.rstd), _ptr(gamma), _ptr(beta), n, c, st.groups,
                d * h * w, 1, _ptr(coef), _ptr(dgamma), _ptr(dbeta), stream_ptr())
        tot = red.view(n, c * 2).sum(0)
        torch.distributed.all_reduce(tot, group=None if st.sync_group is True else st.sync_group)
        red = torch.zeros_like(red)
        red.view(n, c * 2)[0] = tot
        dgamma = dbeta = None
    if dx_sums is not None:
        assert world == 1 and st.sums is not None and dx_sums.numel() == c
        _launch("b200_norm_bwd_finalize_sums", _ptr(red), _ptr(st.mean), _ptr(st.rstd), _ptr(gamma), _ptr(beta), n, c, st.groups,
                d * h * w, 1 if st.batch_stats else 0, _ptr(coef), _ptr(dgamma), _ptr(dbeta), _ptr(st.sums), _ptr(dx_sums), stream_ptr())
    else:
        _launch("b200_norm_bwd_finalize", _ptr(red), _ptr(st.mean), _ptr(st.rstd), _ptr(gamma), _ptr(beta), n, c, st.groups,
                d * h * w * world, 1 if st.batch_stats else 0, _ptr(coef), _ptr(dgamma), _ptr(dbeta), stream_ptr())
    if dx is not None:
        if fast and write_g:
            _launch("b200_norm_bwd_apply_g", _ref(x), _ref(dy), _ptr(coef), _ref(dx), 1 if accumulate else 0, stream_ptr(),
                    shape=x.shape)
        elif fast:
            _launch("b200_norm_silu_bwd_apply_fast", _ref(x), _ref(dy), _ptr(coef), _ref(dx), 1 if accumulate else 0, stream_ptr(),
                    shape=x.shape)
        else:
            _launch("b200_norm_act_bwd_apply", _ref(x), _ref(dy), ACT[act], _ptr(coef), _ref(dx), 1 if accumulate else 0,
                    stream_ptr(), shape=x.shape)


def act_bwd(x, dy, act: str, dx, accumulate=False):
    _launch("b200_act_bwd", _ref(x), _ref(dy), ACT[act], _ref(dx), 1 if accumulate else 0, stream_ptr())


# --------------------------------------------------------------------------------------------- element-wise
OP_ADD, OP_MUL, OP_ADD_RELU, OP_COPY, OP_SIGMOID = 0, 1, 2, 3, 4


def binary(a, b, y, op: int):
    _launch("b200_binary", _ref(a), _ref(b), _ref(y), op, stream_ptr())
    return y


def gate_bwd(x, psi, dout, dpsi, dx, accumulate=False):
    _launch("b200_gate_bwd", _ref(x), _ref(psi), _ref(dout), _ref(dpsi), _ref(dx), 1 if accumulate else 0, stream_ptr())


def relu_mask_bwd(y, dy, da):
    _launch("b200_relu_mask_bwd", _ref(y), _ref(dy), _ref(da), stream_ptr())


def convert(src, dst):
    _launch("b200_convert", _ref(src), _ref(dst), stream_ptr())
    return dst


def softmax_channels(x, y, c0: int, c1: int):
    _launch("b200_softmax_channels", _ref(x), _ref(y), c0, c1, stream_ptr())


# --------------------------------------------------------------------------------------------------- losses
def _check_target(pred, target, per_voxel: int, what: str):
    """The loss kernels index the dense target by the prediction's voxel count: a short buffer would be read out of bounds."""
    vox = pred.shape[0] * pred.shape[1] * pred.shape[2] * pred.shape[3]
    if target.numel() != vox * per_voxel or not target.is_contiguous():
        raise _lib.B200Error(f"{what}: target has {target.numel()} elements (contiguous={target.is_contiguous()}), the prediction "
                             f"{tuple(pred.shape)} needs {vox * per_voxel}")


def bce_logits(logits, target_f32, dlogits=None, grad_scale: float = 1.0) -> torch.Tensor:
    """Returns the SUM of the per-element losses as a 1-element float64 tensor (device)."""
    _check_target(logits, target_f32, logits.shape[-1], "bce_logits")
    out = zeros(1, torch.float64, logits.device)
    _launch("b200_bce_logits", _ref(logits), _ptr(target_f32), _ptr(out), _ref(dlogits), float(grad_scale), stream_ptr())
    return out


def n2v_mse_sums(pred, target_f32) -> torch.Tensor:
    _check_target(pred, target_f32, 2 * pred.shape[-1], "n2v_mse")
    out = zeros(2, torch.float64, pred.device)
    _launch("b200_n2v_mse", _ref(pred), _ptr(target_f32), _ptr(out), None, 1.0, 0, stream_ptr())
    return out


def n2v_mse_bwd(pred, target_f32, dpred, grad_scale: float):
    _check_target(pred, target_f32, 2 * pred.shape[-1], "n2v_mse")
    _launch("b200_n2v_mse", _ref(pred), _ptr(target_f32), None, _ref(dpred), float(grad_scale), 1, stream_ptr())


def n2v_mse_fused(pred, target_f32, dpred, grad_scale: float) -> torch.Tensor:
    """One pass: (sum of squared masked errors, sum of the mask) and dpred = -2 (t - y m) m * grad_scale; the division by the
    mask count is done on the device by the optimiser (`optim_step_dev(denom=sums[1:2])`): no host read of a scalar."""
    _check_target(pred, target_f32, 2 * pred.shape[-1], "n2v_mse")
    out = zeros(2, torch.float64, pred.device)
    _launch("b200_n2v_mse", _ref(pred), _ptr(target_f32), _ptr(out), _ref(dpred), float(grad_scale), 2, stream_ptr())
    return out


def softmax_ce(logits, target_i64, dlogits=None, grad_scale: float = 1.0, ignore_index: int = -100) -> torch.Tensor:
    """float64[3] on the device: (loss sum over the counted voxels, counted voxels, voxels with an illegal label)."""
    _check_target(logits, target_i64, 1, "softmax_ce")
    if target_i64.dtype != torch.int64:
        raise _lib.B200Error(f"softmax_ce: target must be int64 class indices, got {target_i64.dtype}")
    out = zeros(3, torch.float64, logits.device)
    _launch("b200_softmax_ce", _ref(logits), _ptr(target_i64), _ptr(out), _ref(dlogits), float(grad_scale), int(ignore_index),
            stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------ optimiser
def adamw_step(p, g, m, v, lr, beta1, beta2, eps, wd, step: int, grad_scale: float = 1.0):
    _launch("b200_adamw_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), lr, beta1, beta2, eps, wd, step, grad_scale,
            stream_ptr())


def adam_step(p, g, m, v, lr, beta1, beta2, eps, wd, step: int, grad_scale: float = 1.0):
    """torch.optim.Adam: the weight decay is an L2 term of the gradient (TRAIN.OPTIMIZER = 'ADAM')."""
    _launch("b200_adam_step", _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), lr, beta1, beta2, eps, wd, step, grad_scale,
            stream_ptr())


def sgd_step(p, g, mom, lr, momentum, wd, first: bool, grad_scale: float = 1.0, nesterov: bool = False):
    _launch("b200_sgd_step", _ptr(p), _ptr(g), _ptr(mom), p.numel(), lr, momentum, wd, 1 if first else 0, grad_scale,
            1 if nesterov else 0, stream_ptr())


OPT_KIND = {"adamw": 0, "adam": 1, "sgd": 2}
HP_LR, HP_BETA1, HP_BETA2, HP_EPS, HP_WD, HP_MOMENTUM, HP_NESTEROV, HP_GRAD_SCALE, HP_CLIP, HP_SIZE = 0, 1, 2, 3, 4, 5, 6, 7, 8, 16


def optim_step_dev(kind: str, p, g, m, v, hp_dev, state_dev, derived_dev, gsq=None, denom=None):
    """Optimiser update with device-resident hyper-parameters (two launches: a one-thread prepare + the element kernel);
    capturable in a CUDA graph.  See `b200_optim_step_dev` in include/biapy_b200.h for the layouts."""
    global LAUNCHES
    LAUNCHES += 1
    _launch("b200_optim_step_dev", OPT_KIND[kind], _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(hp_dev), _ptr(gsq),
            _ptr(denom), _ptr(state_dev), _ptr(derived_dev), stream_ptr())


def write_floats(dst: torch.Tensor, values):
    """dst[:len(values)] = values in stream order (values are kernel arguments: no staging buffer, no host race)."""
    arr = (C.c_float * len(values))(*values)
    _launch("b200_write_floats", _ptr(dst), arr, len(values), stream_ptr())


def scale_by_dev(g: torch.Tensor, denom: torch.Tensor, mul: float = 1.0):
    _launch("b200_scale_by_dev", _ptr(g), g.numel(), _ptr(denom), float(mul), stream_ptr())


def sumsq(g, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = zeros(1, torch.float64, g.device)
    _launch("b200_sumsq", _ptr(g), g.numel(), _ptr(out), stream_ptr())
    return out
