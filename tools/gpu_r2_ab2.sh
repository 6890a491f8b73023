#!/bin/bash
# A/B: dense gradient hand-over of the max-pool backward, transposed-conv output box width
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_models.py tests/test_gpu_engine.py tests/test_gpu_ops.py tests/test_gpu_workflow.py -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
run() { name=$1; shift; env "$@" timeout 600 python bench.py --detail --dtype bf16 --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"; tail -2 gpurun_out/bench_$name.err; }
run base B200_POOL_DENSE=1 B200_CONVT_BOX=64
run box16 B200_POOL_DENSE=1 B200_CONVT_BOX=16
run box32 B200_POOL_DENSE=1 B200_CONVT_BOX=32
run nodense B200_POOL_DENSE=0 B200_CONVT_BOX=64
python - <<'PY'
import json
for f in ("base", "box16", "box32", "nodense"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_{f}.json") if l.startswith("{")][-1])
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "launches", d["gpu_launches"],
              {k: v["ms_per_step"] for k, v in d["roofline"]["all"].items() if "convT_fprop" in k or "maxpool" in k or k == "binary"})
    except Exception as e:
        print(f, "ERR", e)
PY
