"""CPU check of the plain SIMT kernels (no GPU in the build container).

The kernel *source text* is cut out of ``biapy_b200/csrc/*.cu``, compiled with g++ behind ``tests/simt_emu/emu.h`` (one
std::thread per CUDA thread, barriers for ``__syncthreads`` / ``__shfl_xor_sync``, 2-D grids, dynamic shared memory) and run
against straightforward loops / textbook formulas in double precision:

* the coalesced pointwise convolutions (fprop / dgrad / wgrad of the 1-2 channel layers) and the compile-time-window max-pool;
* the radix-select histogram of the percentile clipping, and the equal-bin histogram behind the Otsu threshold against
  ``np.histogram`` itself (data, edges and counts handed over in a file);
* the three spline overlap-add kernels (plain, cover-mask, slot) against the reference's scatter loop written out in C++, bit for
  bit, on grids with padding, non-monotone starts, triple overlaps, two channels and z slabs;
* the weight pack / unpack kernels against independent statements of their layouts, and the one-launch batched pack kernel
  (``pack_batch.cuh``) against them;
* the loss kernels (BCE with logits, soft-max cross-entropy, Noise2Void masked MSE: sums and gradients) and the fused AdamW / Adam / SGD (+ Nesterov) kernels, by-value and device-hyper-parameter forms, against the torch.optim update rules in double precision;
* the whole GroupNorm / InstanceNorm + activation chain: ``channel_sums`` -> ``norm_finalize`` -> ``scale_shift_act_rows`` and
  ``norm_act_bwd_reduce`` -> ``norm_bwd_finalize`` -> ``norm_act_bwd_apply_rows`` (dx, dgamma, dbeta) on channel slices, and the
  channel sums of dx that ``norm_bwd_finalize`` derives from the reductions (the bias gradient of the convolution in front of the
  normalisation) against the brute-force sum of the reference dx; ``sums_through_pointwise`` against W^T s.

This is test infrastructure: it proves indexing, shuffle patterns, reduction layouts and formulas, not speed; the device run of
the same kernels is covered by the ``-m gpu`` parity tests.  It is what lets a kernel be changed with no GPU at hand.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "biapy_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "simt_emu")

KERNELS = {
    "conv_simt.cu": ["conv1x1_cout_cv_kernel", "conv1x1_cin_cv_kernel", "conv1x1_wgrad_head_cv_kernel",
                     "conv1x1_wgrad_image_cv_kernel", "pack_weight_kernel", "unpack_wgrad_kernel"],
    "ops.cu": ["maxpool_fwd_win_kernel", "maxpool_bwd_win_kernel", "channel_sums_kernel", "norm_finalize_kernel",
               "scale_shift_act_rows_kernel", "norm_act_bwd_reduce_kernel", "norm_bwd_finalize_kernel", "sums_through_pointwise_kernel",
               "norm_act_bwd_apply_rows_kernel", "adamw_kernel", "adam_kernel", "sgd_kernel", "optim_prepare_kernel",
               "optim_dev_kernel", "bce_logits_kernel", "bce_logits_dense_kernel", "n2v_mse_kernel", "softmax_ce_kernel"],
    "ends.cu": ["select_hist_kernel", "edge_hist_kernel"],
    "stitch.cu": ["overlap_add_kernel", "overlap_add_cover_kernel", "overlap_add_slot_kernel"],
    "conv_umma.cu": ["pack_weight_xfold_kernel", "pack_convT_weight_kernel", "unpack_convT_wgrad_kernel"],
    "conv_xline.cu": ["pack_weight_xline_kernel"],
}
# helper definitions that sit right above a kernel and are cut out together with it
PREAMBLE = {
    "select_hist_kernel": "template <typename S> __device__ __forceinline__ uint32_t select_key",
    "pack_weight_xfold_kernel": "__host__ __device__ inline bool xfold_geom",
    "pack_weight_xline_kernel": "__host__ __device__ inline void xline_pack_index",
    "bce_logits_kernel": "__device__ __forceinline__ void block_atomic_add",
    "overlap_add_kernel": "struct MergeParams {",
    "overlap_add_slot_kernel": "struct SlotRow {",
}


def cut_kernel(src: str, name: str) -> str:
    """`[template <...>] __global__ void ... name(...) { ... }` as it stands in the source."""
    m = re.search(r"__global__[^;{]*?\b" + re.escape(name) + r"\(", src)
    assert m, f"kernel {name} not found"
    start = src.rfind("\n", 0, m.start()) + 1
    prev = src.rfind("\n", 0, start - 1) + 1
    if src[prev:start].startswith("template <"):
        start = prev
    # the brace that opens the body: first '{' after the parameter list closes
    depth, j = 0, m.end() - 1
    while True:
        ch = src[j]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                break
        j += 1
    i = src.index("{", j)
    depth, k = 0, i
    while True:
        ch = src[k]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                break
        k += 1
    return src[start:k + 1]


def test_simt_kernels_on_the_host_emulator(tmp_path):
    parts = []
    # activation helpers of common.cuh and the kernel-side tensor view of ops.cu, as they stand
    with open(os.path.join(CSRC, "common.cuh")) as f:
        common = f.read()
    parts.append("// ---- common.cuh: activations\n" + common[common.index("__device__ __forceinline__ float act_fwd(int act, float x)"):
                                                                  common.index("#define B200_DISPATCH_ACT")])
    with open(os.path.join(CSRC, "ops.cu")) as f:
        ops_src = f.read()
    a = ops_src.index("template <typename T>\nstruct View {")
    parts.append("// ---- ops.cu: View\n" + ops_src[a:ops_src.index("};", a) + 2])
    for fname, names in KERNELS.items():
        with open(os.path.join(CSRC, fname)) as f:
            src = f.read()
        for n in names:
            body = cut_kernel(src, n)
            if n in PREAMBLE:
                a = src.index(PREAMBLE[n])
                body = src[a:src.index(body)] + body
            # dynamic shared memory: `extern __shared__ T name[];` -> the emulator's per-block buffer
            body = re.sub(r"extern __shared__ ([\w ]+?) (\w+)\[\];", r"\1* \2 = reinterpret_cast<\1*>(g_ctx->dyn_smem);", body)
            parts.append(f"// ---- {fname}: {n}\n" + body)
    # the batched pack kernel is held to the single-job kernels it replaces inside a training pass
    with open(os.path.join(CSRC, "pack_batch.cuh")) as f:
        staged = f.read()
    parts.append("// ---- pack_batch.cuh\n" + staged[staged.index("enum PackKind"):])
    with open(os.path.join(CSRC, "norm_fast.cuh")) as f:
        staged = f.read()
    staged = re.sub(r"extern __shared__ ([\w ]+?) (\w+)\[\];", r"\1* \2 = reinterpret_cast<\1*>(g_ctx->dyn_smem);", staged)
    parts.append("// ---- norm_fast.cuh\n" + staged[staged.index("__device__ __forceinline__ float tanh_approx"):])
    (tmp_path / "kernels.inc").write_text("\n\n".join(parts) + "\n")
    # np.histogram's answer for the Otsu histogram kernel (numpy is the reference there): data, edges, counts in one binary file
    import numpy as np
    from oracle import port_norm
    blobs = []
    for name, img in port_norm.otsu_cases().items():
        flat = img.reshape(-1)
        if np.all(flat == flat[0]):
            continue
        counts, edges = np.histogram(flat, bins=256)
        assert edges.dtype == np.float32
        blobs.append(np.int64(flat.size).tobytes() + flat.tobytes() + edges.tobytes() + counts.astype(np.int64).tobytes())
    (tmp_path / "hist_cases.bin").write_bytes(np.int64(len(blobs)).tobytes() + b"".join(blobs))
    exe = tmp_path / "simt_emu"
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-I", str(tmp_path), "-I", EMU, "-I", os.path.join(ROOT, "include"),
           os.path.join(EMU, "driver.cpp"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    try:
        r = subprocess.run([str(exe), str(tmp_path / "hist_cases.bin")], capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        pytest.fail("emulated kernels dead-locked (a shuffle or barrier not reached by every thread)")
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]
