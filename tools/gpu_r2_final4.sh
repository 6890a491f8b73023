#!/bin/bash
# Round-2 closing visit (third session): whole parity suite, bench lines, launch list, ncu captures of the x-line weight gradient
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_golden.jsonl
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep "\[smoke\]" gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --detail --dtype bf16 --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_detail_bf16.json 2> gpurun_out/bench_detail_bf16.err; echo "bench detail rc=$?"
timeout 900 python bench.py --detail --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_detail_fp16.json 2> gpurun_out/bench_detail_fp16.err; echo "bench detail fp16 rc=$?"
timeout 600 python bench.py --workload infer --steps 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"; tail -3 gpurun_out/bench_infer.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_detail_bf16", "bench_detail_fp16", "bench_infer", "bench_reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "frac", r.get("frac"), r.get("kernel"),
              "launches", d.get("gpu_launches"), "skipped", d.get("skipped_steps"), "other", d.get("other_dtype"), "dtype", d.get("dtype"))
        if r.get("members"):
            print("   members", r["members"])
        if d.get("infer"):
            i = d["infer"]
            print("  infer", round(i["value"], 1), "e2e", round(i["e2e"]["value"], 1), "merge", i["roofline"]["achieved"], i["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_r2d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/ncu_bench.log 2>&1
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_wgrad_xline -c 4 \
  -o gpurun_out/xline_wgrad_step_r2d -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/ncu_xline_wgrad_step.log 2>&1
ls -la gpurun_out | tail -8
