#!/bin/bash
# 2-GPU visit: NCCL checks of the data-parallel paths + the weak-scaling bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\|Error\|error" gpurun_out/dist_check.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_2gpu.json") if l.startswith("{")][-1])
    print("2 GPUs:", d["ms_per_step"], d["value"], d["e2e"]["value"], d["n_gpus"])
except Exception as e:
    print("ERR", e)
PY
tail -5 gpurun_out/bench_2gpu.err
