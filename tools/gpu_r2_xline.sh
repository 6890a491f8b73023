#!/bin/bash
# x-line visit: op tests, training / inference bench with the x-line kernels off and on, ncu --set full of the x-line launches
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_xline.py -q -s 2>&1 | grep -E "xline vs|labels|passed|failed" | cut -c1-400
for m in 0 1; do
  B200_XLINE=$m timeout 600 python bench.py --no-cpu-baseline --no-infer > gpurun_out/bench_xline$m.json 2> gpurun_out/bench_xline$m.err; echo "bench XLINE=$m rc=$?"
done
B200_XLINE=1 B200_XLINE_FUSE=1 timeout 600 python bench.py --no-cpu-baseline --no-infer > gpurun_out/bench_xline1_fuse1.json 2> gpurun_out/bench_xline1_fuse1.err
for m in 0 1; do
  B200_XLINE=$m timeout 600 python bench.py --workload infer --steps 3 --no-cpu-baseline > gpurun_out/bench_infer_xline$m.json 2> gpurun_out/bench_infer_xline$m.err; echo "infer XLINE=$m rc=$?"
done
B200_XLINE=1 B200_XLINE_FUSE=1 timeout 600 python bench.py --workload infer --steps 3 --no-cpu-baseline > gpurun_out/bench_infer_xline1_fuse1.json 2> gpurun_out/bench_infer_xline1_fuse1.err
python - <<'PY'
import json
for f in ("bench_xline0", "bench_xline1", "bench_xline1_fuse1", "bench_infer_xline0", "bench_infer_xline1", "bench_infer_xline1_fuse1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "frac", r.get("frac"), r.get("kernel"),
              "launches", d.get("gpu_launches"), "other", d.get("other_dtype"), "dtype", d.get("dtype"))
        a = r.get("all") or {}
        print("   ", {k: v for k, v in a.items() if "xline" in k})
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fprop_xline -c 10 -o gpurun_out/xline_r2 -f python tools/xline_ncu.py > gpurun_out/ncu_xline.log 2>&1
ls -la gpurun_out | tail -12
