#!/bin/bash
# Round-2 visit 2: batched packs, slot merge v2, Cout slices / Cin 512 transposed conv, Otsu; bench lines again.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_baseline_configs.py::test_full_size_cfg1_resunet128 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -q -s -k cfg1 > gpurun_out/pytest_cfg1.log 2>&1; echo "cfg1 rc=$?"; grep "cfg1\|passed\|failed\|Error" gpurun_out/pytest_cfg1.log | cut -c1-400
for v in "0 2" "16 2" "32 2" "64 2" "128 2" "32 1" "32 4"; do set -- $v; B200_MERGE_ROWS=$1 B200_MERGE_UNROLL=$2 timeout 300 python tools/merge_micro.py 2>&1 | tail -1; done | tee gpurun_out/merge_micro.log
timeout 900 python bench.py --detail --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --workload cfg4 --no-cpu-baseline --steps 10 --detail > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 rc=$?"; tail -3 gpurun_out/bench_cfg4.err
timeout 600 python bench.py --workload cfg3 --no-cpu-baseline --detail > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "cfg3 rc=$?"; tail -3 gpurun_out/bench_cfg3.err
python - <<'PY'
import json
for f in ("bench", "bench_cfg3", "bench_cfg4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", r.get("frac"), r.get("kernel"),
              "launches", d["gpu_launches"], "skipped", d.get("skipped_steps"), "other", d.get("other_dtype"))
        if d.get("infer"):
            i = d["infer"]
            print("  infer", round(i["value"], 1), "e2e", round(i["e2e"]["value"], 1), "merge", i["roofline"]["achieved"], i["roofline"]["frac"])
        for k, v in list(r.get("all", {}).items())[:(60 if f == "bench" else 25)]:
            print("   ", k, v)
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:overlap_add -c 1 -o gpurun_out/merge_r2b -f python tools/merge_micro.py > gpurun_out/ncu_merge.log 2>&1
ls -la gpurun_out | tail -12
