#!/bin/bash
# x-line weight gradient: parity tests and op timings (x-folded next to x-line)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_xline.py -q -k wgrad -s 2>&1 | grep -E "xline wgrad|passed|failed|Error|error" | cut -c1-300
timeout 300 python tools/xline_probe.py --quick --wgrad 2>&1 | grep -E "time|OK|FAILED" | cut -c1-200
