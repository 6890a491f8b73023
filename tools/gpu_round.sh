#!/bin/bash
# One GPU-box visit: parity tests, bench lines (training step, sliding-window inference), ncu launch list of the timed region,
# one full capture of the top kernel family.
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --detail > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "bench infer rc=$?"; cut -c1-600 gpurun_out/bench_infer.json
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_fprop_xslab -c 3 \
  -o gpurun_out/top_kernel -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python - <<'PY'
import json
for f in ("bench", "bench_infer"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la gpurun_out
