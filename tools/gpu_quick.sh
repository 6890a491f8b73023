#!/bin/bash
# quick visit: smoke + parity tests + the two bench lines (no profiler passes)
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload infer > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "bench infer rc=$?"; tail -3 gpurun_out/bench_infer.err
python - <<'PY'
import json
for f in ("bench", "bench_infer"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, d["ms_per_step"], d["value"], d["e2e"], {k: r.get(k) for k in ("kernel", "achieved", "peak", "frac", "traffic")}, d["cpu_baseline"], d["clocks"], d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
