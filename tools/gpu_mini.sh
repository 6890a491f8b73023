#!/bin/bash
# short visit: the pooling / pointwise op tests and the model-level parity tests, then one training bench line
set -x
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -q -x -k "maxpool or pointwise or golden_forward" > gpurun_out/pytest_mini.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_mini.log | cut -c1-200
timeout 60 python bench.py --no-cpu-baseline > gpurun_out/bench_idx32.json 2> gpurun_out/bench_idx32.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_idx32.json').read().strip().splitlines()[-1]); print('bench_idx32', d['ms_per_step'], d['value'])"
