#!/bin/bash
# memcheck of the kernels written in round 2 (slot / slab merge, patch-range crop, Otsu histogram, batched packs, device-hp optimiser,
# fast norm chain, streaming tiles) + the whole suite once more + the bench line with the pipelined upload
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stitch.py tests/test_gpu_ends.py tests/test_gpu_ops.py tests/test_gpu_chunks.py -q -x -k "not full_size and not golden_3d" > gpurun_out/sanitize.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/sanitize.log | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_engine.py -q -x -k "adam or fp16 or cross_entropy or pinned" > gpurun_out/sanitize_engine.log 2>&1; echo "memcheck engine rc=$?"; tail -8 gpurun_out/sanitize_engine.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench.json") if l.startswith("{")][-1])
i = d["infer"]
print("train", d["ms_per_step"], d["value"], d["e2e"]["value"], "infer", i["ms_per_volume"], i["value"], "e2e", i["e2e"]["value"], i["roofline"]["frac"])
PY
