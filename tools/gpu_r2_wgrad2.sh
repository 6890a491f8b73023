#!/bin/bash
# x-line weight gradient, second visit: step with the kernel on both shapes, ncu --set full of the 16 -> 16 launch
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --detail --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_xwgrad2_detail.json 2> gpurun_out/bench_xwgrad2_detail.err
python - <<'PY'
import json
for f in ("bench_xwgrad2_detail",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "frac", r.get("frac"),
              "launches", d.get("gpu_launches"), "dtype", d.get("dtype"))
        a = r.get("all") or {}
        for k, v in sorted(a.items(), key=lambda kv: -kv[1]["ms_per_step"])[:32]:
            print("   ", k, v)
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_xline -c 2 -o gpurun_out/xline_wgrad_r2_v4 -f python tools/xline_wgrad_ncu.py > gpurun_out/ncu_xline_wgrad_v4.log 2>&1
tail -2 gpurun_out/ncu_xline_wgrad_v4.log
