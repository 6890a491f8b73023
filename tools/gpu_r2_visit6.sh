#!/bin/bash
# analytic bias gradients (normalisation backward reductions): parity suites, then the step with the switch off / on
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_engine.py tests/test_gpu_xline.py tests/test_gpu_baseline_configs.py tests/test_gpu_workflow.py tests/test_gpu_ops.py -q -x 2>&1 | tail -5 | cut -c1-400
for m in 0 1; do
B200_DBIAS_ANALYTIC=$m timeout 600 python bench.py --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_v6_analytic$m.json 2> gpurun_out/bench_v6_analytic$m.err
done
python - <<'PY'
import json
for f in ("bench_v6_analytic0", "bench_v6_analytic1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "frac", r.get("frac"),
              "launches", d.get("gpu_launches"), "dtype", d.get("dtype"))
    except Exception as e:
        print(f, "ERR", e)
PY
