#!/bin/bash
# memset instead of torch fill kernels, output-indexed weight packs: engine tests, step, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_engine.py tests/test_gpu_workflow.py -q -x 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-infer > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_v9.json") if l.startswith("{")][-1])
print("ms", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "other", d.get("other_dtype"), d["clocks"])
a = d["roofline"].get("all") or {}
for k in ("pack_batch", "optim_step_dev", "bce_logits", "norm_finalize", "norm_bwd_finalize_sums", "sums_through_pointwise"):
    print("  ", k, a.get(k))
PY
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_r2e.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2e.csv gpurun_out/launches_r2e_summary.csv; grep -E "pack_batch|elementwise|fill|memset|sumsq|finalize|bce" gpurun_out/launches_r2e_summary.csv | head -12
