#!/bin/bash
# Round-2 visit 4: A/B of the norm backward forms and the merge occupancy variant, golden parity numbers of the three engines
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_golden.jsonl
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_ops.py tests/test_gpu_stitch.py -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
for v in "5" "6" "5" "6"; do B200_MERGE_OCC=$v timeout 300 python tools/merge_micro.py 2>&1 | tail -1; done | tee gpurun_out/merge_micro.log
for v in g recompute; do
  B200_NORM_BWD=$v timeout 600 python bench.py --detail --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_bwd_$v.json 2> gpurun_out/bench_bwd_$v.err; echo "bench $v rc=$?"
done
B200_NORM_FAST=1 timeout 600 python bench.py --dtype fp16 --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_fp16_fast.json 2> gpurun_out/bench_fp16_fast.err; echo "bench fp16 fast rc=$?"
python - <<'PY'
import json
for f in ("bench_bwd_g", "bench_bwd_recompute", "bench_fp16_fast"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1))
        for k, v in list(r.get("all", {}).items()):
            if "norm" in k or "silu" in k or "pack" in k:
                print("   ", k, v)
    except Exception as e:
        print(f, "ERR", e)
PY
