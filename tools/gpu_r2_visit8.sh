#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_engine.py tests/test_gpu_ops.py tests/test_gpu_baseline_configs.py tests/test_gpu_workflow.py -q -x 2>&1 | tail -4 | cut -c1-300
bash tools/gpu_r2_ab.sh B200_DBIAS_ANALYTIC
