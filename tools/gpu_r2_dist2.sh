#!/bin/bash
# 2-GPU visit of the closing build: NCCL checks + the bench line at N = 2, and the reference arm under torchrun (rank 0 only)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\|Error\|error\|Traceback" gpurun_out/dist_check.log | head -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_2gpu.json") if l.startswith("{")][-1])
    print("2 GPUs train:", d["ms_per_step"], d["value"], d["e2e"]["value"], d["n_gpus"])
    i = d["infer"]
    print("2 GPUs infer:", i["ms_per_volume"], i["value"], i["e2e"]["value"], i["config"]["patches_this_rank"], i["config"]["exchange_bytes_received_rank0"])
except Exception as e:
    print("ERR", e)
PY
tail -5 gpurun_out/bench_2gpu.err
