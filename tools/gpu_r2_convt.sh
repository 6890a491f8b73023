#!/bin/bash
# transposed-conv TMA-store epilogue: parity tests, A/B against the direct scatter epilogue
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_models.py tests/test_gpu_engine.py tests/test_gpu_baseline_configs.py -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
for v in 64 32 16 0; do
  if [ $v = 0 ]; then export B200_CONVT_TMA=0; else export B200_CONVT_TMA=1 B200_CONVT_BOX=$v; fi
  timeout 600 python bench.py --detail --dtype bf16 --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_convt_$v.json 2> gpurun_out/bench_convt_$v.err; echo "bench convt $v rc=$?"; tail -2 gpurun_out/bench_convt_$v.err
done
unset B200_CONVT_TMA B200_CONVT_BOX
python - <<'PY'
import json
for f in ("bench_convt_64", "bench_convt_32", "bench_convt_16", "bench_convt_0"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), {k: v for k, v in d["roofline"]["all"].items() if "convT" in k})
    except Exception as e:
        print(f, "ERR", e)
PY
