#!/bin/bash
# Round-2 closing visit (third session, final build): whole parity suite, smoke, the fp16 step with the backward SiLU chain exact / one-MUFU,
# the bench line, the inference line
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_golden.jsonl
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep "\[smoke\]" gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
for m in bf16 auto; do
B200_NORM_FAST=$m timeout 600 python bench.py --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_normfast_$m.json 2> gpurun_out/bench_normfast_$m.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_normfast_$m.json") if l.startswith("{")][-1])
print("NORM_FAST=$m ms", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["clocks"])
PY
done
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench.json") if l.startswith("{")][-1])
r = d.get("roofline") or {}
print("bench ms", round(d["ms_per_step"], 3), "value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "frac", r.get("frac"), "launches", d.get("gpu_launches"),
      "other", d.get("other_dtype"), "cpu", d["cpu_baseline"]["value"], d["config"].get("allreduce"))
i = d["infer"]
print("  infer", round(i["value"], 1), "e2e", round(i["e2e"]["value"], 1), "merge", i["roofline"]["achieved"], i["roofline"]["frac"])
PY
