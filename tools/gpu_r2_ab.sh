#!/bin/bash
# step A/B of one environment switch: tools/gpu_r2_ab.sh VAR [extra bench args]
set -x
mkdir -p gpurun_out
V=$1; shift
for m in 0 1 0 1; do
env $V=$m timeout 600 python bench.py --no-cpu-baseline --no-infer --no-other-dtype "$@" > gpurun_out/bench_ab_${V}_$m.json 2> gpurun_out/bench_ab_${V}_$m.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_ab_${V}_$m.json") if l.startswith("{")][-1])
    print("$V=$m", "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d.get("gpu_launches"), d.get("clocks"))
except Exception as e:
    print("$V=$m ERR", e)
PY
done
