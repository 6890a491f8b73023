#!/usr/bin/env python
"""Headline benchmark: 3D patches/s of the 128^3 x 2ch Residual U-Net training step (BASELINE config[1]) on N B200s, with the
sliding-window inference of config[2] (ONE 512^3 volume sharded over the N ranks) in the same JSON line.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port), rank 0 only
    python bench.py --workload cfg3 | cfg4 | infer           # other BASELINE configs: Attention U-Net N2V fp16, 2D U-Net bf16, inference only

A training step = forward + loss + backward + gradient all-reduce + optimiser on one batch of synthetic patches per GPU (weak
scaling).  `value` is measured with the batch resident in HBM; `e2e` runs the same K steps through the reference-facing call
``train_one_epoch`` (``biapy/engine/train_engine.py:25``) over a loader of pinned host batches: H2D copies inside the timed
region, the loss read back every step.  Prints ONE JSON line on rank 0.
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
FWD_GFLOP_PER_PATCH = 305.61      # conv + convT, 2*MAC
STEP_GFLOP_PER_PATCH = 913.08     # fwd + dgrad + wgrad minus dgrad of the two input-fed layers

# BASELINE.json configs as training workloads: (model class, kwargs, per-GPU batch shape, target channels, loss, dtype, label)
WORKLOADS = {
    # default engine dtype fp16: the 16-bit storage format whose outputs meet the 1e-3 tolerance at full size (profiles/
    # parity_cfg1_r2.json: 6.1e-4; bf16, the label BASELINE carries, 5.1e-3) -- same tcgen05 kind::f16 kernels, same MMA rate;
    # the bf16 engine is measured in the same run and reported under "other_dtype"
    "train": dict(arch="resunet", kw=CFG2, batch=(BATCH, 128, 128, 128, 2), tgt_c=1, loss="bce", dtype="fp16", ndim=3,
                  unit="patches/s", metric="3D patches/sec (128^3x2ch bf16 ResU-Net training step)",
                  label="BASELINE config[1]: 3D Residual U-Net fm[16,32,64,128,256] gn/silu, 128^3x2ch, batch 4 per GPU, "
                        "training step = fwd + BCEWithLogits + bwd + grad all-reduce + AdamW(lr 1e-3, wd 0.02)",
                  step_gflop=STEP_GFLOP_PER_PATCH * BATCH),
    "cfg3": dict(arch="attention_unet",
                 kw=dict(image_shape=(64, 64, 64, 1), activation="elu", feature_maps=[16, 32, 64, 128, 256], drop_values=[0] * 5,
                         normalization="in", k_size=3, yx_down=[2] * 4, z_down=[2] * 4, isotropy=[True] * 5, larger_io=False,
                         conv_layers=[2] * 5, output_channels=[1]),
                 batch=(8, 64, 64, 64, 1), tgt_c=2, loss="n2v_mse", dtype="fp16", ndim=3, unit="patches/s",
                 metric="3D patches/sec (64^3x1ch fp16 Attention U-Net denoising training step)",
                 label="BASELINE config[3]: 3D Attention U-Net fm[16..256] in/elu, 64^3x1ch, batch 8 per GPU, Noise2Void masked MSE "
                       "(mask density 0.198 %), grad all-reduce, AdamW", step_gflop=None),
    "cfg4": dict(arch="unet",
                 kw=dict(image_shape=(512, 512, 3), activation="elu", feature_maps=[32, 64, 128, 256, 512], drop_values=[0] * 5,
                         normalization="in", k_size=3, yx_down=[2] * 4, z_down=[2] * 4, isotropy=[True] * 5, larger_io=False,
                         conv_layers=[2] * 5, output_channels=[2]),
                 batch=(32, 512, 512, 3), tgt_c=2, loss="bce", dtype="bf16", ndim=2, unit="images/s",
                 metric="2D images/sec (512x512x3 bf16 U-Net training step)",
                 label="BASELINE config[4]: 2D U-Net fm[32..512] in/elu, 512x512x3, 2 output channels, batch 32 per GPU, "
                       "BCEWithLogits, grad all-reduce, AdamW", step_gflop=None),
}
DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}


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


def model_class(arch):
    from biapy_b200.models.attention_unet import Attention_U_Net
    from biapy_b200.models.resunet import ResUNet
    from biapy_b200.models.unet import U_Net
    return {"resunet": ResUNet, "unet": U_Net, "attention_unet": Attention_U_Net}[arch]


def build_model(dtype, arch="resunet", kw=CFG2):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = model_class(arch)(**kw)
    return m.cuda().set_engine(dtype=dtype)


def synth_batch(seed, spec):
    """Synthetic fp16 patches + target in BiaPy's channels-last layout (N, [Z,] Y, X, C)."""
    g = torch.Generator().manual_seed(seed)
    shape = tuple(spec["batch"])
    x = torch.randn(*shape, generator=g).to(torch.float16)
    tshape = shape[:-1] + (spec["tgt_c"],)
    if spec["loss"] == "n2v_mse":                      # target || mask, mask density of 3d_denoising.yaml:11
        t = torch.randn(*tshape, generator=g)
        half = spec["tgt_c"] // 2
        t[..., half:] = (torch.rand(*(shape[:-1] + (half,)), generator=g) < 0.00198).float()
        t = t.to(torch.float16)
    else:
        t = (torch.rand(*tshape, generator=g) < 0.3).to(torch.float16)  # Bernoulli(0.3) mask
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


# ----------------------------------------------------------------------------------------------------------- CPU arm
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


def host_cpus():
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return {"os_cpu_count": os.cpu_count(), "affinity": avail}


def pick_cpu_threads():
    """Thread count for the CPU arm = "all the host threads it can use".  os.cpu_count() is not it: the GPU boxes report 128 CPUs
    but run this container under a smaller CPU quota (measured there: 16 threads 0.20 s, 32: 0.36 s, 64: 0.88 s, 128: 38 s per
    64^3 step), so the best of a short calibration on a 64^3 patch is used."""
    avail = host_cpus()["affinity"]
    cands = sorted({c for c in (8, 16, 32, 64) if c <= avail} | {min(avail, 8)})
    best, best_t, table = cands[0], None, {}
    for c in cands:
        cpu_step_time(c, patch=64)
        t = cpu_step_time(c, patch=64)
        table[c] = round(t, 3)
        if best_t is None or t < best_t:
            best, best_t = c, t
        elif t > 2.5 * best_t:
            break                                   # past the quota: more threads only thrash
    return best, table


def cpu_baseline_train(all_threads=None):
    """Bounded sample on the host cores: one 128^3 training step with all usable threads AND with 4 threads -- what BiaPy itself
    would use (main_threads = min(4, cpus_per_rank), biapy/utils/misc.py:1216-1263, applied at _biapy.py:340)."""
    calib = None
    if all_threads is None:
        all_threads, calib = pick_cpu_threads()
    t_all = cpu_step_time(all_threads)
    t_4 = cpu_step_time(4)
    return {"value": 1.0 / t_all, "unit": "patches/s", "cores": all_threads, "kind": "port",
            "sample": "one training step on one 128^3x2 patch (batch 1), fp32, torch CPU via oracle/port_models.py",
            "threads_all": {"threads": all_threads, "patches_per_s": 1.0 / t_all, "s_per_step": t_all},
            "threads_4": {"threads": 4, "patches_per_s": 1.0 / t_4, "s_per_step": t_4,
                          "why": "BiaPy's own setting: min(4, cpus_per_rank) (misc.py:1216-1263)"},
            "host": host_cpus(), "calibration_64cube_s_per_step": calib}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    spec = WORKLOADS["train"]
    threads, calib = pick_cpu_threads()
    torch.set_num_threads(threads)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_step_time(threads)
    ts = [cpu_step_time(threads) for _ in range(max(1, args.steps))]
    sec = sum(ts) / len(ts)
    v = 1.0 / sec
    t4 = cpu_step_time(4)
    out = {"impl": "reference", "metric": spec["metric"], "value": v, "unit": "patches/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": spec["label"], "sample": "bounded sample: one 128^3x2 patch (batch 1) per step on the host cores, fp32 "
                                                           "(the reference has no AMP); the B200 arm runs batch 4 per GPU"},
           "cpu_baseline": {"value": v, "unit": "patches/s", "cores": threads, "kind": "port",
                            "sample": "one 128^3x2 patch per step, fwd+bwd+AdamW, torch CPU fp32 via oracle/port_models.py",
                            "threads_all": {"threads": threads, "patches_per_s": v, "s_per_step": sec},
                            "threads_4": {"threads": 4, "patches_per_s": 1.0 / t4, "s_per_step": t4,
                                          "why": "BiaPy's own setting: min(4, cpus_per_rank) (misc.py:1216-1263)"},
                            "host": host_cpus(), "calibration_64cube_s_per_step": calib},
           "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------- training arm
def _trainer_for(spec, dtype):
    from biapy_b200.engine.train import Trainer
    model = build_model(dtype, spec["arch"], spec["kw"])
    return Trainer(model, loss=spec["loss"], optimizer="adamw", lr=1e-3, weight_decay=0.02)


def _epoch_cfg(spec):
    from biapy_b200.config.config import load_config
    return load_config({"PROBLEM": {"NDIM": "3D" if spec["ndim"] == 3 else "2D"},
                        "DATA": {"PATCH_SIZE": tuple(spec["kw"]["image_shape"])}})


def measure_training(spec, dtype_name, args, rank, world, local, with_profile=True, with_e2e=True):
    """Device-resident K steps (CUDA-graph replay + all-reduce + optimiser), then K steps end to end through train_one_epoch."""
    from biapy_b200 import ops
    from biapy_b200.engine.train_engine import train_one_epoch
    trainer = _trainer_for(spec, DTYPES[dtype_name])
    xh, th = synth_batch(100 + rank, spec)
    xh, th = xh.pin_memory(), th.pin_memory()
    xd, td = xh.cuda(), th.cuda()

    def dev_step(i):
        trainer.step(xd, td)

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
    if not args.graph and with_profile:      # eager mode: per-kernel CUDA events inside the timed region itself
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
    launches = ops.LAUNCHES - n0
    res = {"ms": ms, "launches": launches, "clocks": clocks, "prof": prof, "region": "timed region"}
    if with_e2e:
        # the reference-facing call: one train_one_epoch over a loader of K pinned host batches (H2D inside, loss read back per step)
        cfg = _epoch_cfg(spec)
        loader = [(xh, th)] * args.steps
        warm = [(xh, th)] * min(args.warmup, 2)

        def epoch(batches):
            with contextlib.redirect_stdout(io.StringIO()):
                stats, _ = train_one_epoch(cfg, trainer.model, None, None, None, lambda t, b: t, batches, [trainer], "cuda", 0)
            return stats

        epoch(warm)
        stats_box = {}
        ms_e2e = timed(lambda i: stats_box.update(epoch(loader)), 1, world) / args.steps
        res.update(ms_e2e=ms_e2e, final_loss=float(stats_box.get("loss", float("nan"))),
                   h2d=xh.numel() * xh.element_size() + th.numel() * th.element_size())
    if args.graph and with_profile:
        # kernels inside a replayed graph cannot carry events: time them in an eager pass of the same steps right after
        g, trainer._graph = trainer._graph, None
        ops.PROFILE = {} if rank == 0 else None
        ops.PROFILE_SHAPES = args.detail
        timed(dev_step, args.steps, world)
        res["prof"], ops.PROFILE = ops.PROFILE, None
        trainer._graph = g
        res["region"] = "separate eager pass of the same steps in this process (timed region replays a CUDA graph)"
    res["skipped_steps"] = int(trainer._opt_state[1].item())
    res["allreduce"] = ("none (one rank)" if world == 1 else
                        f"flat fp32 gradient; elements [{trainer._split[1]}, {trainer.fp.grad.numel()}) reduced on a side stream under the second "
                        f"CUDA graph of the pass (backward steps below {trainer._split[0]}), the rest after it"
                        if getattr(trainer, "_graph_b", None) is not None else "one all-reduce of the flat fp32 gradient after the pass")
    del trainer
    torch.cuda.empty_cache()
    return res


def roofline_from_profile(prof, steps, peaks, timed_region_s, region):
    agg = {}
    for name, recs in prof.items():
        t = sum(a.elapsed_time(b) for a, b, _, _ in recs)
        agg[name] = (t, sum(r[2] for r in recs), sum(r[3] for r in recs), len(recs))
    # dominant family among the tensor-bound ones (conv1x1_* / conv_image_* are HBM streams: see ops._family).  The forward /
    # input-gradient convolutions of the small-channel layers run on two kernel families since round 2 (x-folded: conv_umma.cu,
    # x-line: conv_xline.cu, chosen per layer by engine/tape.py): they are ONE family for the roofline -- the same layers, the
    # same algorithmic FLOPs -- with the per-kernel split kept in `all` and `members`.
    def group_of(k):
        fam = k.split(" ")[0]
        return "conv_fprop" if fam.startswith("conv_fprop_") else fam
    tensor_keys = [k for k in agg if agg[k][1] > 0 and not k.startswith(("conv1x1_", "conv_image_"))]
    groups = {}
    for k in tensor_keys or [k for k in agg if agg[k][1] > 0]:
        g = groups.setdefault(group_of(k), [0.0, 0.0, 0.0, 0, {}])
        g[0] += agg[k][0]; g[1] += agg[k][1]; g[2] += agg[k][2]; g[3] += agg[k][3]
        fam = k.split(" ")[0]
        m = g[4].setdefault(fam, [0.0, 0.0, 0])
        m[0] += agg[k][0]; m[1] += agg[k][1]; m[2] += agg[k][3]
    top = max(groups, key=lambda g: groups[g][0])
    t, fl, by, cnt, members = groups[top]
    members = {k: {"ms_per_step": round(v[0] / steps, 3), "TFLOP/s": round(v[1] / (v[0] / 1e3) / 1e12, 1), "launches": v[2] // steps}
               for k, v in sorted(members.items(), key=lambda kv: -kv[1][0])}
    ach = fl / (t / 1e3) / 1e12
    # a short timed region runs at boost clocks: the burst figure is the like-for-like denominator; long runs settle at the sustained one
    burst = timed_region_s < 2.0
    peak = peaks["tf_burst"] if burst else peaks["tf_sust"]
    traffic, traffic_note = None, None
    for f in ("traffic_r2.json", "traffic_r1.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", f)))
            fam = next((m for m in members if m in tj), None)      # ncu capture of the group's most expensive kernel family
            if fam is not None:
                traffic = tj[fam]["dram_bytes_per_launch"]
                traffic_note = dict(tj[fam], family=fam, other_captures={m: tj[m] for m in members if m in tj and m != fam})
                break
        except (OSError, ValueError, KeyError):
            pass
    if len(members) > 1:
        top = top + " (" + " + ".join(members) + ")"
    total_flops = sum(v[1] for v in agg.values())
    return {"kernel": top, "region": region, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": ach / peak, "frac_of_sustained": ach / peaks["tf_sust"], "frac_of_burst": ach / peaks["tf_burst"],
            "traffic": traffic, "traffic_note": traffic_note, "launches": cnt, "ms_in_step": t / steps, "members": members,
            "peak_source": peaks["src"] + (", burst figure (timed region %.2f s, boost clocks)" % timed_region_s if burst
                                           else ", sustained figure (kernel timed inside a %.1f s region)" % timed_region_s),
            "algorithmic_gflop_per_step_from_events": total_flops / steps / 1e9,
            "all": {k: {"ms_per_step": round(v[0] / steps, 3), "TFLOP/s": round(v[1] / (v[0] / 1e3) / 1e12, 1),
                        "launches": v[3] // steps} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}}


def run_train(args):
    rank, world, local = dist_setup(args.gpus)
    peaks = load_peaks()
    spec = WORKLOADS[args.workload]
    dtype_name = args.dtype or spec["dtype"]
    main = measure_training(spec, dtype_name, args, rank, world, local)
    other = None
    if args.workload == "train" and args.other_dtype:
        alt = "fp16" if dtype_name == "bf16" else "bf16"
        o = measure_training(spec, alt, args, rank, world, local, with_profile=False, with_e2e=False)
        other = {"dtype": alt, "ms_per_step": o["ms"], "value": spec["batch"][0] * world / (o["ms"] / 1e3),
                 "skipped_steps": o["skipped_steps"]}
    infer = None
    if args.workload == "train" and args.infer:
        infer = measure_inference(args, rank, world, local, peaks, DTYPES[dtype_name], steps=max(2, min(args.steps, 3)))
    if rank != 0:
        return
    n_units = spec["batch"][0] * world
    ms = main["ms"]
    roof = roofline_from_profile(main["prof"], args.steps, peaks, ms * args.steps / 1e3, main["region"]) if main["prof"] else None
    cpu = cpu_baseline_train() if (args.cpu_baseline and world == 1 and args.workload == "train") else None
    step_gflop = spec["step_gflop"] or (roof["algorithmic_gflop_per_step_from_events"] if roof else None)
    parity = None
    try:
        parity = json.load(open(os.path.join(ROOT, "profiles", "parity_cfg1_r2e.json")))
    except (OSError, ValueError):
        pass
    out = {
        "metric": spec["metric"], "value": n_units / (ms / 1e3), "unit": spec["unit"],
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype_name, "data": "synthetic",
        "config": {"workload": spec["label"],
                   "global_batch": n_units, "parallelism": f"dp{world}", "cuda_graph": bool(args.graph),
                   "allreduce": main.get("allreduce"),
                   "l2": "per-step working set (activations + gradients, several GB) >> 126 MB L2; no explicit flush",
                   "algorithmic_gflop_per_step": step_gflop,
                   "dtype_note": "fp16 storage / fp32 accumulation (tcgen05 kind::f16): the 16-bit engine whose outputs meet the 1e-3 "
                                 "tolerance; BASELINE's bf16 label is the other_dtype line (same kernels, bf16 storage)"
                                 if dtype_name == "fp16" else None,
                   "parity": {"tolerance": "1e-3 rel (north_star)", "full_size_table": "profiles/parity_cfg1_r2e.json "
                              "(tests/test_gpu_baseline_configs.py::test_full_size_cfg1_resunet128)",
                              "meets_1e-3_on_outputs": ["float32", "float16"], "measured": parity}},
        "e2e": {"value": n_units / (main["ms_e2e"] / 1e3), "unit": spec["unit"], "ms_per_step": main["ms_e2e"],
                "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": 8,
                "api": "biapy_b200.engine.train_engine.train_one_epoch(cfg, model, ..., loader of pinned host fp16 batches, [trainer]) "
                       "-- the reference's epoch call (biapy/engine/train_engine.py:25); loss read back every step (8 bytes, lagged)"},
        "gpu_launches": main["launches"], "step_tflops": (step_gflop / ms) if step_gflop else None,
        "final_loss": main.get("final_loss"), "skipped_steps": main["skipped_steps"],
        "clocks": main["clocks"], "roofline": roof, "cpu_baseline": cpu, "other_dtype": other, "infer": infer,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------ inference arm
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


def measure_inference(args, rank, world, local, peaks, dtype, steps, with_cpu=False):
    """BASELINE config[2]: ONE 512^3 volume, 128^3 patches, 25 % overlap -> 216 patches, sigmoid head, sharded over the `world`
    ranks (strong scaling): every rank crops / predicts its patch range and merges the output slab it owns; the volume is left
    sharded (each rank returns its planes), as a by-chunks writer would store it."""
    from biapy_b200 import ops
    from biapy_b200.data import _stitch
    from biapy_b200.engine.inference import predict_volume, shard_planes
    model = build_model(dtype).eval()
    V, P = 512, 128
    # every rank generates only the planes its patches read (same generator stream per plane block -> the same volume everywhere)
    kw = dict(overlap=(0.25,) * 3, padding=(0, 0, 0), batch_size=4, head_activations=["ce_sigmoid"])
    a, b = shard_planes((V, V, V, 2), (P, P, P, 2), kw["overlap"], kw["padding"], "reflect", rank, world)
    vol_h = torch.empty(b - a, V, V, 2, dtype=torch.float16).pin_memory()
    for z in range(a, b):
        vol_h[z - a] = torch.randn(V, V, 2, generator=torch.Generator().manual_seed(1000 + z)).to(torch.float16)
    shard_d = _stitch.VolumeShard(vol_h.cuda(), a, V)
    pkw = dict(kw, rank=rank, world=world, gather="none")
    stats = {}

    def step(i):
        predict_volume(model, shard_d, (P, P, P, 2), stats=stats, **pkw)

    z0, z1 = (V * rank) // world, (V * (rank + 1)) // world
    out_h = torch.empty(z1 - z0, V, V, 1, dtype=torch.float32).pin_memory()

    def e2e_step(i):                                          # host volume planes in, host prediction slab out
        r = predict_volume(model, _stitch.VolumeShard(vol_h, a, V), (P, P, P, 2), **pkw)      # pinned planes: upload pipelined with the batches
        slab = r[0] if isinstance(r, tuple) else r
        out_h.copy_(slab, non_blocking=True)

    step(0)
    sampler = ClockSampler(local)
    sampler.start()
    n0 = ops.LAUNCHES
    ms = timed(step, steps, world)
    launches = ops.LAUNCHES - n0
    clocks = sampler.stop()
    e2e_step(0)
    ms_e2e = timed(e2e_step, max(2, steps), world)
    del model
    # roofline of the HBM-bound stitch kernel: spline overlap-add of 216 fp32 patch predictions into the volume
    roof = None
    if rank == 0:
        axes = [_stitch.Axis(V, P, 0, 0.25) for _ in range(3)]
        pred = torch.rand(216, P, P, P, 1, device="cuda")
        starts, wins = [ax.starts(1) for ax in axes], [ax.window() for ax in axes]
        merge = lambda i: _stitch.merge_device(pred, (V, V, V), starts, wins, (0, 0, 0))
        merge(0)
        ms_merge = timed(merge, 10, 1)
        alg = 216 * P ** 3 * 4 + V ** 3 * 4                  # every patch element read once + every output element written once
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2.json")))["overlap_add"]["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        roof = {"kernel": "overlap_add_slot (spline-weighted merge of 216 x 128^3 fp32 patches into 512^3, one GPU)", "bound": "hbm",
                "achieved": alg / (ms_merge / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": alg / (ms_merge / 1e3) / 1e9 / peaks["hbm"], "traffic": traffic, "ms_per_launch": ms_merge,
                "algorithmic_bytes": alg, "peak_source": peaks["src"]}
        del pred
    torch.cuda.empty_cache()
    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        return None
    cpu = None
    if with_cpu:
        cores, _ = pick_cpu_threads()
        sec, t_fwd, t_st = cpu_infer_patch_time(cores)
        cpu = {"value": 1.0 / sec, "unit": "patches/s", "cores": cores, "kind": "port",
               "sample": f"eval forward + sigmoid of one 128^3x2 patch ({t_fwd:.2f} s) + numpy crop / spline merge of a 256^3 volume "
                         f"(27 patches, {t_st:.2f} s per patch), fp32, oracle port"}
    return {"metric": "3D patches/sec (sliding-window inference, ONE 512^3 volume over all ranks)", "value": 216 / (ms / 1e3),
            "unit": "patches/s", "n_gpus": world, "steps": steps, "ms_per_volume": ms, "scaling": "strong",
            "dtype": str(dtype).replace("torch.", ""),
            "config": {"workload": "BASELINE config[2]: 512^3x2 fp16 volume -> 216 patches of 128^3 (25 % overlap) -> ResUNet forward -> "
                                   "sigmoid -> spline overlap-add; ranks own contiguous patch ranges and z slabs of the output, exchange "
                                   "only the patch pieces that reach into a neighbour's slab; output left sharded",
                       "patches_this_rank": stats.get("patches"), "exchange_bytes_received_rank0": stats.get("exchange_bytes_received"),
                       "l2": "per-rank patches + predictions (hundreds of MB) >> 126 MB L2; no explicit flush"},
            "e2e": {"value": 216 / (ms_e2e / 1e3), "unit": "patches/s", "ms_per_volume": ms_e2e,
                    "h2d_bytes_per_step": vol_h.numel() * 2, "d2h_bytes_per_step": out_h.numel() * 4,
                    "api": "biapy_b200.engine.inference.predict_volume(VolumeShard(pinned host fp16 planes), rank, world, gather='none') "
                           "-> pinned host fp32 slab (bytes are rank 0's; the upload is pipelined with the forward passes)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}


def run_infer(args):
    rank, world, local = dist_setup(args.gpus)
    peaks = load_peaks()
    dtype_name = args.dtype or "bf16"
    r = measure_inference(args, rank, world, local, peaks, DTYPES[dtype_name], steps=max(2, args.steps),
                          with_cpu=args.cpu_baseline and world == 1)
    if rank != 0:
        return
    r.update(steps=args.steps, warmup=args.warmup, ms_per_step=r["ms_per_volume"], higher_is_better=True, vs_baseline=None,
             data="synthetic")
    print(json.dumps(r), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "cfg3", "cfg4", "infer"])
    ap.add_argument("--dtype", default=None, choices=["bf16", "fp16", "fp32"], help="engine dtype (default: the workload's own)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-infer", dest="infer", action="store_false", help="skip the config[2] inference sub-object")
    ap.add_argument("--no-other-dtype", dest="other_dtype", action="store_false", help="skip the second 16-bit dtype line")
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
