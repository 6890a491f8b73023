#!/bin/bash
# Round-2 visit 1: smoke, parity suite (incl. the full-size config[1] table), bench lines, A/B of the new kernels, ncu.
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep "\[smoke\]" gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_baseline_configs.py::test_full_size_cfg1_resunet128 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -q -s -k cfg1 > gpurun_out/pytest_cfg1.log 2>&1; echo "cfg1 rc=$?"; grep "cfg1\|passed\|failed\|Error" gpurun_out/pytest_cfg1.log | cut -c1-400
for v in "slot 1" "slot 2" "slot 4" "cover 2"; do set -- $v; B200_MERGE_KERNEL=$1 B200_MERGE_ROWS=$2 timeout 300 python tools/merge_micro.py 2>&1 | tail -1; done | tee gpurun_out/merge_micro.log
timeout 900 python bench.py --detail > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
B200_NORM_FAST=0 timeout 600 python bench.py --detail --no-infer --no-other-dtype --no-cpu-baseline > gpurun_out/bench_nofast.json 2> gpurun_out/bench_nofast.err; echo "bench nofast rc=$?"
timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "cfg3 rc=$?"; tail -3 gpurun_out/bench_cfg3.err
timeout 600 python bench.py --workload cfg4 --no-cpu-baseline --steps 10 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 rc=$?"; tail -3 gpurun_out/bench_cfg4.err
python - <<'PY'
import json
for f in ("bench", "bench_nofast", "bench_cfg3", "bench_cfg4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", r.get("frac"),
              "launches", d["gpu_launches"], "skipped", d.get("skipped_steps"), "other", d.get("other_dtype"))
        if d.get("infer"):
            i = d["infer"]
            print("  infer", round(i["value"], 1), "e2e", round(i["e2e"]["value"], 1), "merge", i["roofline"]["achieved"], i["roofline"]["frac"])
        if f in ("bench", "bench_nofast"):
            for k, v in list(r.get("all", {}).items())[:45]:
                print("   ", k, v)
    except Exception as e:
        print(f, "ERR", e)
PY
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:overlap_add -c 2 -o gpurun_out/merge_r2 -f python tools/merge_micro.py > gpurun_out/ncu_merge.log 2>&1
ls -la gpurun_out | tail -30
