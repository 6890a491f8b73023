#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -q -s > gpurun_out/pytest_cfg.log 2>&1; echo "pytest rc=$?"; grep "^\[\|passed\|failed\|Error\|assert" gpurun_out/pytest_cfg.log | cut -c1-250 | tail -20
