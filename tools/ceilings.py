"""Per-kernel ceilings of the cfg-2 training step next to what `bench.py --detail` measured.

    python tools/ceilings.py profiles/bench_detail_r1_step15p2ms.json > profiles/ceilings_r1.md

For every labelled launch group of the detail file: the algorithmic work (SURVEY 8d conventions: 2 * MAC for convolutions,
one read / write per operand for the streams), the bound that applies (tensor pipe at the measured sustained bf16 rate; for the
x-folded N = 4 * Cout kernels also the shared-memory operand-fetch bound of DESIGN 3.0, max(N/2, (128 + N)/4) cycles per
M = 128, K = 16 MMA including the structural zeros of the Toeplitz fold; HBM at the measured copy rate), the time that bound
allows, and measured / ceiling.  Host arithmetic only; peaks from MEASURED_PEAKS.json.
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BATCH = 4
ELT = 2          # bf16


# launch groups of the cfg-2 model that the detail file labels without shapes: transposed convolutions 256->256 @8^3->16^3,
# 128->128 @16^3, 64->64 @32^3, 32->32 @64^3->128^3 (k = s = 2) and the four max-pools behind the encoder levels
_CONVT = [(256, 8), (128, 16), (64, 32), (32, 64)]          # (channels, input edge)
_POOL = [(16, 128), (32, 64), (64, 32), (128, 16)]          # (channels, input edge)


def _convt(reads_in, reads_out, writes_in, writes_out):
    def f(hbm, tf):
        flop = sum(2.0 * BATCH * e ** 3 * c * c * 8 for c, e in _CONVT)
        byts = sum(BATCH * e ** 3 * c * ELT * (reads_in + writes_in) + BATCH * (2 * e) ** 3 * c * ELT * (reads_out + writes_out)
                   for c, e in _CONVT)
        t = max(flop / tf, byts / hbm)
        return f"{flop / 1e9:.1f} GF, {byts / 1e6:.0f} MB", "hbm" if byts / hbm > flop / tf else "tensor", t * 1e3
    return f


def _pool(passes_fine, passes_coarse):
    def f(hbm, tf):
        byts = sum(BATCH * e ** 3 * c * ELT * passes_fine + BATCH * (e // 2) ** 3 * c * ELT * passes_coarse for c, e in _POOL)
        return f"{byts / 1e6:.0f} MB", "hbm", byts / hbm * 1e3
    return f


FIXED = {
    "convT_fprop_tc": _convt(1, 0, 0, 1),            # read x, write the fine tensor
    "convT_dgrad_tc": _convt(0, 1, 1, 0),            # read the fine gradient, write dx
    "convT_wgrad_tc": _convt(1, 1, 0, 0),            # read both
    "maxpool_fwd": _pool(1, 1),                      # read x, write y
    "maxpool_bwd": _pool(3, 1),                      # read x, read + write dx (accumulate into the skip gradient), read dy
}


def main(path):
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"] * 1e9
    tf = peaks["bf16_tflops_sustained"] * 1e12
    clk = peaks["sm_max_mhz"] * 1e6
    sms = 148
    d = json.loads(open(path).read().strip().splitlines()[-1])
    rows = []
    for label, v in d["roofline"]["all"].items():
        ms, n = v["ms_per_step"], v["launches"]
        m = re.match(r"(\w+) (\d+)->(\d+) k(\d)(\d)(\d) @(\d+)x(\d+)x(\d+)", label)
        bound = work = ceil_ms = None
        if m:
            fam, cin, cout = m.group(1), int(m.group(2)), int(m.group(3))
            kd, kh, kw = int(m.group(4)), int(m.group(5)), int(m.group(6))
            vox = BATCH * int(m.group(7)) * int(m.group(8)) * int(m.group(9))
            flop = 2.0 * vox * cin * cout * kd * kh * kw * n
            byts = vox * (cin + cout) * ELT * n
            t_tensor, t_hbm = flop / tf, byts / hbm
            work = f"{flop / 1e9:.1f} GF, {byts / 1e6:.0f} MB"
            if "xfold" in fam and kw == 3 and cout * 4 <= 256 and cin % 16 == 0:
                # x-fold: rows of 4 voxels, N = 4*Cout, K per (dz, dy) = (3 + kw) * Cin, MMA of K = 16
                N = 4 * cout
                mmas = (vox / 4 / 128) * kd * kh * ((3 + kw) * cin / 16) * n
                t_smem = mmas * max(N / 2, (128 + N) / 4) / (sms * clk)
                cands = {"tensor": t_tensor, "hbm": t_hbm, "smem (x-fold)": t_smem}
            else:
                cands = {"tensor": t_tensor, "hbm": t_hbm}
            bound = max(cands, key=cands.get)
            ceil_ms = cands[bound] * 1e3
        else:
            m = re.match(r"(\w+) c(\d+) @(\d+)x(\d+)x(\d+)", label)
            if m:
                fam, c = m.group(1), int(m.group(2))
                el = BATCH * int(m.group(3)) * int(m.group(4)) * int(m.group(5)) * c * n
                per = {"channel_sums": 1, "scale_shift_act": 2, "norm_act_bwd_reduce": 2, "norm_act_bwd_apply": 3,
                       "scale_shift_silu_fast": 2, "norm_silu_bwd_reduce_g": 2, "norm_bwd_apply_g": 3,
                       "norm_silu_bwd_apply_fast": 3}.get(fam)
                if per:
                    byts = el * per * ELT
                    bound, ceil_ms, work = "hbm", byts / hbm * 1e3, f"{byts / 1e6:.0f} MB"
        if bound is None and label in FIXED:
            work, bound, ceil_ms = FIXED[label](hbm, tf)
        rows.append((ms, label, n, work, bound, ceil_ms))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    known = [r for r in rows if r[5] is not None]
    print(f"# Ceilings of the cfg-2 training step ({os.path.basename(path)})\n")
    print(f"Peaks: {peaks['bf16_tflops_sustained']} TFLOP/s sustained bf16, {peaks['hbm_gbs']} GB/s copy, {peaks['sm_max_mhz']:.0f} MHz "
          f"(MEASURED_PEAKS.json).  Sum of the eager per-launch events: {total:.2f} ms (graph replay: {d['ms_per_step']:.2f} ms).\n")
    print("| launch group | launches | measured ms | algorithmic work | bound | ceiling ms | measured / ceiling |")
    print("|---|---|---|---|---|---|---|")
    for ms, label, n, work, bound, ceil_ms in rows:
        if ms < 0.03:
            continue
        if ceil_ms is None:
            print(f"| {label} | {n} | {ms:.3f} | | | | |")
        else:
            print(f"| {label} | {n} | {ms:.3f} | {work} | {bound} | {ceil_ms:.3f} | {ms / ceil_ms:.1f}x |")
    km, kc = sum(r[0] for r in known), sum(r[5] for r in known)
    print(f"\nGroups with a modelled ceiling: {km:.2f} ms measured against {kc:.2f} ms of ceilings ({km / kc:.1f}x); "
          f"the other {total - km:.2f} ms are weight packing / unpacking and coefficient kernels (5-10 us launches), strided copies, "
          f"the loss and the optimiser.")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "bench_detail_r1_step15p2ms.json"))
