#!/bin/bash
set -x
mkdir -p gpurun_out
# (1) N = 256 x-slab launch under ncu, outside a graph
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_fprop_x -c 8 --csv --log-file gpurun_out/diag_n256.csv \
  python -m pytest tests/test_gpu_umma.py -x -q -k "test_conv_fprop_xfold and case7" > gpurun_out/diag_n256.log 2>&1; echo rc=$?
# (2) eager bench under the duration-only pass
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_eager.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_eager.log 2>&1; echo rc=$?
tail -3 gpurun_out/launches_eager.csv | cut -c1-200
# (3) graph bench, cache control off
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches_graph_nocc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_graph_nocc.log 2>&1; echo rc=$?
tail -3 gpurun_out/launches_graph_nocc.csv | cut -c1-200
