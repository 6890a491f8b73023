#!/bin/bash
# x-line weight-gradient visit: op tests, step A/B with the kernel off / on, ncu --set full of the 16 -> 16 launch
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_xline.py -q 2>&1 | tail -3 | cut -c1-300
for m in 0 1; do
  B200_XLINE_WGRAD=$m timeout 600 python bench.py --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_xwgrad$m.json 2> gpurun_out/bench_xwgrad$m.err; echo "bench XLINE_WGRAD=$m rc=$?"
done
B200_XLINE_WGRAD=1 timeout 600 python bench.py --detail --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_xwgrad1_detail.json 2> gpurun_out/bench_xwgrad1_detail.err
python - <<'PY'
import json
for f in ("bench_xwgrad0", "bench_xwgrad1", "bench_xwgrad1_detail"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "frac", r.get("frac"),
              "launches", d.get("gpu_launches"), "dtype", d.get("dtype"))
        a = r.get("all") or {}
        for k, v in a.items():
            if "wgrad" in k and "128x128x128" in k:
                print("   ", k, v)
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_xline -c 2 -o gpurun_out/xline_wgrad_r2 -f python tools/xline_wgrad_ncu.py > gpurun_out/ncu_xline_wgrad.log 2>&1
tail -3 gpurun_out/ncu_xline_wgrad.log
ls -la gpurun_out | tail -6
