#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_engine.py -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep "^\[\|passed\|failed\|Error" gpurun_out/pytest_gpu.log | cut -c1-250 | tail -20
timeout 600 python bench.py --detail --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_q.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["value"], d["final_loss"])
for k, v in d["roofline"]["all"].items():
    if "norm_act_bwd" in k and v["ms_per_step"] > 0.05: print(k, v["ms_per_step"])
PY
