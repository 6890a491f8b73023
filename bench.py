#!/usr/bin/env python
"""Headline benchmark: 3D patches/s of the 128^3 x 2ch bf16 Residual U-Net training step (BASELINE config[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port), rank 0 only
    python bench.py --workload infer                         # BASELINE config[2]: 512^3 sliding-window inference

A step = forward + BCEWithLogits + backward + gradient all-reduce + AdamW on one batch of 4 synthetic patches per
GPU (weak scaling).  `value` is measured with the batch resident in HBM; `e2e` runs the same step through the
public Trainer API from pinned host fp16 buffers (H2D inside the timed region, loss read back every step).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

CFG2 = dict(image_shape=(128, 128, 128, 2), activation="silu", feature_maps=[16, 32, 64, 128, 256], drop_values=[0] * 5,
            normalization="gn", k_size=3, yx_down=[2] * 4, z_down=[2] * 4, isotropy=[True] * 5, larger_io=False,
            conv_layers=[2] * 5, output_channels=[1])
BATCH = 4
METRIC = "3D patches/sec (128^3x2ch bf16 ResU-Net training step)"
WORKLOAD = ("BASELINE config[1]: 3D Residual U-Net fm[16,32,64,128,256] gn/silu, 128^3x2ch, batch 4 per GPU, "
            "training step = fwd + BCEWithLogits + bwd + grad all-reduce + AdamW(lr 1e-3, wd 0.02)")
FWD_GFLOP_PER_PATCH = 305.61      # conv + convT, 2*MAC
STEP_GFLOP_PER_PATCH = 913.08     # fwd + dgrad + wgrad minus dgrad of the two input-fed layers


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_model(dtype):
    from biapy_b200.models.resunet import ResUNet
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ResUNet(**CFG2)
    return m.cuda().set_engine(dtype=dtype)


def synth_batch(seed, n=BATCH):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 128, 128, 128, 2, generator=g).to(torch.float16)         # synthetic fp16 volume patches
    t = (torch.rand(n, 128, 128, 128, 1, generator=g) < 0.3).to(torch.float16)  # Bernoulli(0.3) mask
    return x, t


def dist_setup(n_gpus):
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def timed(fn, steps, world):
    """barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks -> ms per step"""
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.barrier()
        ms = t.item()
    return ms / steps


def cpu_step_time(threads, reps=1, patch=128):
    """One training step (fwd + BCE + bwd + AdamW) of the reference path on the host cores, batch 1, fp32,
    through the oracle port (the reference's own Python cannot travel to this box)."""
    from oracle import port_models
    from biapy_b200.models.resunet import ResUNet
    torch.set_num_threads(threads)
    kw = dict(CFG2)
    kw["image_shape"] = (patch, patch, patch, 2)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ResUNet(**kw)
    sd = {k: v.clone().requires_grad_(True) for k, v in m.state_dict().items()}
    opt = torch.optim.AdamW(list(sd.values()), lr=1e-3, weight_decay=0.02)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 2, patch, patch, patch, generator=g)
    t = (torch.rand(1, 1, patch, patch, patch, generator=g) < 0.3).float()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        y = port_models.forward("resunet", sd, x, training=True, **kw)
        loss = port_models.bce_with_logits_loss(y, t)
        opt.zero_grad()
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    return min(times)


def pick_cpu_threads():
    """Thread count for the CPU arm.  os.cpu_count() is not it: the GPU boxes report 128 CPUs but run this container under a
    smaller CPU quota (measured there: 16 threads 0.20 s, 32: 0.36 s, 64: 0.88 s, 128: 38 s per 64^3 step), so the best
    of a short calibration on a 64^3 patch is used -- "all the host threads it can use"."""
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cands = sorted({c for c in (8, 16, 32) if c <= avail} | {min(avail, 8)})
    best, best_t = cands[0], None
    for c in cands:
        cpu_step_time(c, patch=64)
        t = cpu_step_time(c, patch=64)
        if best_t is None or t < best_t:
            best, best_t = c, t
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = pick_cpu_threads()
    torch.set_num_threads(threads)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_step_time(threads)
    ts = [cpu_step_time(threads) for _ in range(max(1, args.steps))]
    sec = sum(ts) / len(ts)
    v = 1.0 / sec
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "patches/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "sample": "bounded sample: one 128^3x2 patch (batch 1) per step on the host cores, fp32 "
                                                       "(the reference has no AMP); the B200 arm runs batch 4 per GPU"},
           "cpu_baseline": {"value": v, "unit": "patches/s", "cores": threads, "kind": "port",
                            "sample": "one 128^3x2 patch per step, fwd+bwd+AdamW, torch CPU fp32 via oracle/port_models.py"},
           "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_train(args):
    from biapy_b200 import ops
    from biapy_b200.engine.train import Trainer
    rank, world, local = dist_setup(args.gpus)
    peaks = load_peaks()
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    model = build_model(dtype)
    trainer = Trainer(model, loss="bce", optimizer="adamw", lr=1e-3, weight_decay=0.02)
    xh, th = synth_batch(100 + rank)
    xh, th = xh.pin_memory(), th.pin_memory()
    xd, td = xh.cuda(), th.cuda()
    loss_host = torch.zeros(1, dtype=torch.float64).pin_memory()

    def dev_step(i):
        trainer.step(xd, td)

    def e2e_step(i):
        loss = trainer.step(xh, th)                      # pinned host -> device copies happen inside
        loss_host.copy_(loss, non_blocking=True)

    for i in range(args.warmup):
        dev_step(i)
    torch.cuda.synchronize()
    if args.graph:
        trainer.enable_cuda_graph(xd, td)
        for i in range(2):
            dev_step(i)
        torch.cuda.synchronize()
    n0 = ops.LAUNCHES
    sampler = ClockSampler(local)
    sampler.start()
    if not args.graph:                      # eager mode: per-kernel CUDA events inside the timed region itself
        ops.PROFILE = {} if rank == 0 else None
        ops.PROFILE_SHAPES = args.detail
    ncu_range = os.environ.get("BENCH_PROFILER_RANGE") == "1"   # `ncu --profile-from-start off`: only the timed region
    if ncu_range:
        torch.cuda.profiler.start()
    ms = timed(dev_step, args.steps, world)
    if ncu_range:
        torch.cuda.profiler.stop()
    prof, ops.PROFILE = ops.PROFILE, None
    clocks = sampler.stop()
    launches = (ops.LAUNCHES - n0)
    for i in range(min(args.warmup, 2)):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps, world)
    torch.cuda.synchronize()
    final_loss = float(loss_host.item())
    roofline_region = "timed region"
    if args.graph:
        # kernels inside a replayed graph cannot carry events: time them in an eager pass of the same steps right after
        g, trainer._graph = trainer._graph, None
        ops.PROFILE = {} if rank == 0 else None
        ops.PROFILE_SHAPES = args.detail
        timed(dev_step, args.steps, world)
        prof, ops.PROFILE = ops.PROFILE, None
        trainer._graph = g
        roofline_region = "separate eager pass of the same steps in this process (timed region replays a CUDA graph)"

    if rank != 0:
        return
    patches = BATCH * world
    value = patches / (ms / 1e3)
    # ---- roofline of the dominant kernel class (by summed device time inside the timed region)
    roof = None
    if prof:
        agg = {}
        for name, recs in prof.items():
            t = sum(a.elapsed_time(b) for a, b, _, _ in recs)
            agg[name] = (t, sum(r[2] for r in recs), sum(r[3] for r in recs), len(recs))
        top = max((k for k in agg if agg[k][1] > 0), key=lambda k: agg[k][0])
        t, fl, by, cnt = agg[top]
        ach = fl / (t / 1e3) / 1e12
        # DRAM traffic of the dominant launch of that family, from the committed `ncu --set full` capture (profiles/)
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_r1.json")))
            fam = top.split(" ")[0]               # --detail labels carry the layer shape after the family name
            if fam in tj:
                traffic = tj[fam]["dram_bytes_per_launch"]
                traffic_note = tj[fam]
        except (OSError, ValueError, KeyError):
            pass
        roof = {"kernel": top, "region": roofline_region, "bound": "tensor", "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                "frac": ach / peaks["tf_sust"], "traffic": traffic, "traffic_note": traffic_note, "launches": cnt,
                "ms_in_step": t / args.steps,
                "peak_source": peaks["src"] + ", sustained figure (kernel timed inside a long step)",
                "all": {k: {"ms_per_step": round(v[0] / args.steps, 3), "TFLOP/s": round(v[1] / (v[0] / 1e3) / 1e12, 1),
                            "launches": v[3] // args.steps} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}}
    cpu = None
    if args.cpu_baseline and world == 1:
        cores = pick_cpu_threads()
        sec = cpu_step_time(cores)
        cpu = {"value": 1.0 / sec, "unit": "patches/s", "cores": cores, "kind": "port",
               "sample": "one training step on one 128^3x2 patch (batch 1), fp32, torch CPU via oracle/port_models.py"}
    out = {
        "metric": METRIC, "value": value, "unit": "patches/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch": patches, "parallelism": f"dp{world}", "cuda_graph": bool(args.graph),
                   "l2": "per-step working set (activations + gradients, several GB) >> 126 MB L2; no explicit flush",
                   "algorithmic_gflop_per_step": STEP_GFLOP_PER_PATCH * BATCH},
        "e2e": {"value": patches / (ms_e2e / 1e3), "unit": "patches/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": xh.numel() * xh.element_size() + th.numel() * th.element_size(), "d2h_bytes_per_step": 8,
                "api": "biapy_b200.engine.train.Trainer.step(host fp16 batch, host fp16 target)"},
        "gpu_launches": launches, "step_tflops": STEP_GFLOP_PER_PATCH * BATCH / ms, "final_loss": final_loss,
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(out), flush=True)


def cpu_infer_patch_time(threads):
    """Reference CPU path of the sliding-window pipeline per patch, on a bounded sample: eval forward + sigmoid of ONE 128^3 x 2
    patch (fp32, oracle port) plus numpy crop + spline merge of a 256^3 volume (27 patches) scaled to one patch."""
    from oracle import port_models, port_stitch
    from biapy_b200.models.resunet import ResUNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ResUNet(**CFG2)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn(1, 2, 128, 128, 128, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        port_models.forward("resunet", sd, x[:, :, :64, :64, :64], training=False, **dict(CFG2, image_shape=(64, 64, 64, 2)))   # warm-up
        t0 = time.perf_counter()
        y = port_models.forward("resunet", sd, x, training=False, **CFG2)
        torch.sigmoid(y)
        t_fwd = time.perf_counter() - t0
    vol = np.random.default_rng(0).standard_normal((256, 256, 256, 2)).astype(np.float32)
    t0 = time.perf_counter()
    patches, _ = port_stitch.crop_3d(vol, (128, 128, 128, 2), (0.25,) * 3, (0, 0, 0), "reflect")
    port_stitch.merge_3d(np.ascontiguousarray(patches[..., :1]), (256, 256, 256, 1), (0.25,) * 3, (0, 0, 0))
    t_stitch = (time.perf_counter() - t0) / patches.shape[0]
    return t_fwd + t_stitch, t_fwd, t_stitch


def run_infer(args):
    """BASELINE config[2]: 512^3 volume, 128^3 patches, 25% overlap -> 216 patches, sigmoid head, device-resident."""
    from biapy_b200 import ops
    from biapy_b200.data import _stitch
    from biapy_b200.engine.inference import predict_volume
    rank, world, local = dist_setup(args.gpus)
    peaks = load_peaks()
    model = build_model(torch.bfloat16).eval()
    g = torch.Generator().manual_seed(1)
    vol_h = torch.randn(512, 512, 512, 2, generator=g).to(torch.float16).pin_memory()
    vol_d = vol_h.cuda()
    kw = dict(overlap=(0.25,) * 3, padding=(0, 0, 0), batch_size=4, head_activations=["ce_sigmoid"])

    def step(i):
        predict_volume(model, vol_d, (128, 128, 128, 2), **kw)

    out_h = torch.empty(512, 512, 512, 1, dtype=torch.float32).pin_memory()

    def e2e_step(i):                                          # host volume in, host prediction out
        out_h.copy_(predict_volume(model, vol_h.cuda(non_blocking=True), (128, 128, 128, 2), **kw), non_blocking=True)

    for i in range(max(1, args.warmup // 2)):
        step(i)
    sampler = ClockSampler(local)
    sampler.start()
    n0 = ops.LAUNCHES
    ms = timed(step, args.steps, world)
    launches = ops.LAUNCHES - n0
    clocks = sampler.stop()
    e2e_step(0)
    ms_e2e = timed(e2e_step, max(2, args.steps // 3), world)
    if rank != 0:
        return
    # roofline of the HBM-bound stitch kernels: spline overlap-add of 216 fp32 patch predictions into the volume
    axes = [_stitch.Axis(512, 128, 0, 0.25) for _ in range(3)]
    pred = torch.rand(216, 128, 128, 128, 1, device="cuda")
    starts, wins = [a.starts(1) for a in axes], [a.window() for a in axes]
    merge = lambda i: _stitch.merge_device(pred, (512, 512, 512), starts, wins, (0, 0, 0))
    merge(0)
    ms_merge = timed(merge, 10, 1)
    alg = 216 * 128 ** 3 * 4 + 512 ** 3 * 4                  # every patch element read once + every output element written once
    roof = {"kernel": "overlap_add (spline-weighted merge of 216 x 128^3 fp32 patches into 512^3)", "bound": "hbm",
            "achieved": alg / (ms_merge / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s", "frac": alg / (ms_merge / 1e3) / 1e9 / peaks["hbm"],
            "traffic": None, "ms_per_launch": ms_merge, "algorithmic_bytes": alg, "peak_source": peaks["src"],
            "share_of_step": ms_merge / ms}
    cpu = None
    if args.cpu_baseline and world == 1:
        cores = pick_cpu_threads()
        sec, t_fwd, t_st = cpu_infer_patch_time(cores)
        cpu = {"value": 1.0 / sec, "unit": "patches/s", "cores": cores, "kind": "port",
               "sample": f"eval forward + sigmoid of one 128^3x2 patch ({t_fwd:.2f} s) + numpy crop / spline merge of a 256^3 volume "
                         f"(27 patches, {t_st:.2f} s per patch), fp32, oracle port"}
    print(json.dumps({"metric": "3D patches/sec (sliding-window inference, 512^3 volume)", "value": 216 * world / (ms / 1e3),
                      "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": "BASELINE config[2]: crop 512^3 -> 216x128^3 (25% overlap) -> ResUNet fwd -> "
                                             "sigmoid -> spline overlap-add, one volume per GPU",
                                 "l2": "2 GB volume + 3.6 GB of patches + 1.8 GB of predictions per step >> 126 MB L2; no explicit flush"},
                      "e2e": {"value": 216 * world / (ms_e2e / 1e3), "unit": "patches/s", "ms_per_step": ms_e2e,
                              "h2d_bytes_per_step": vol_h.numel() * 2, "d2h_bytes_per_step": out_h.numel() * 4,
                              "api": "biapy_b200.engine.inference.predict_volume(host fp16 volume) -> host fp32 prediction"},
                      "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "infer"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--detail", action="store_true", help="per-layer-shape kernel table in roofline.all")
    ap.add_argument("--graph", dest="graph", action="store_true", default=True,
                    help="replay forward+backward from a CUDA graph (default)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="eager launches (per-kernel events in the timed region)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (biapy_b200 has no CPU path); use --impl reference for the CPU arm")
    if args.workload == "infer":
        return run_infer(args)
    return run_train(args)


if __name__ == "__main__":
    try:
        main()
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
