#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-graph > gpurun_out/bench_q_eager.json 2> gpurun_out/bench_q_eager.err; echo "bench eager rc=$?"
python - <<'PY'
import json
for f in ("bench_q", "bench_q_eager"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["final_loss"], d["roofline"]["frac"], d["roofline"]["achieved"], d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/bench_q.err
