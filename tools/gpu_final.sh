#!/bin/bash
# one visit, in priority order: full GPU parity suite, bench with the new CUDA-core kernels, the same bench with them switched
# off (A/B), inference bench, smoke.  Every piece writes its own file under gpurun_out/ so a cut-off call still leaves results.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -rf --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log | cut -c1-240
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
B200_POOL_WIN=0 B200_PW_COALESCED=0 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_oldsimt.json 2> gpurun_out/bench_oldsimt.err; echo "bench old rc=$?"
timeout 120 python bench.py --detail --no-cpu-baseline > gpurun_out/bench_detail.json 2> gpurun_out/bench_detail.err; echo "bench detail rc=$?"
timeout 200 python bench.py --workload infer > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "bench infer rc=$?"
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_oldsimt", "bench_detail", "bench_infer"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d.get("e2e", {}).get("value"), d["roofline"]["frac"] if d.get("roofline") else None)
    except Exception as e:
        print(f, "unreadable:", e)
PY
