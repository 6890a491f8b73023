#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of the timed region, one full capture of the top kernel.
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --detail > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
BENCH_PROFILER_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_fprop_xslab -c 3 \
  -o gpurun_out/top_kernel -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
