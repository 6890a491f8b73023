#!/bin/bash
# Round-2 visit 3: slot merge v3, pack_batch block distribution
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stitch.py tests/test_gpu_ops.py tests/test_gpu_engine.py -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
for v in "0 2" "32 2" "64 2" "64 1" "64 4" "0 2"; do set -- $v; B200_MERGE_ROWS=$1 B200_MERGE_UNROLL=$2 timeout 300 python tools/merge_micro.py 2>&1 | tail -1; done | tee gpurun_out/merge_micro.log
timeout 900 python bench.py --detail --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("bench",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", r.get("frac"), r.get("kernel"),
              "launches", d["gpu_launches"], "skipped", d.get("skipped_steps"), "other", d.get("other_dtype"))
        if d.get("infer"):
            i = d["infer"]
            print("  infer", round(i["value"], 1), "e2e", round(i["e2e"]["value"], 1), "merge", i["roofline"]["achieved"], i["roofline"]["frac"])
        for k, v in list(r.get("all", {}).items())[:30]:
            print("   ", k, v)
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:overlap_add -c 1 -o gpurun_out/merge_r2c -f python tools/merge_micro.py > gpurun_out/ncu_merge.log 2>&1
