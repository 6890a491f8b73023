#!/bin/bash
# 2-GPU visit: the captured step with the all-reduce under the second graph -- NCCL checks, then the bench line with the overlap on / off
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?"
grep "dist_check\|Error\|error\|Traceback" gpurun_out/dist_check.log | head -30
for m in 1 0; do
B200_OVERLAP_ALLREDUCE=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-infer --no-other-dtype > gpurun_out/bench_2gpu_ov$m.json 2> gpurun_out/bench_2gpu_ov$m.err; echo "bench2 rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_2gpu_ov$m.json") if l.startswith("{")][-1])
    print("overlap $m: 2 GPUs train:", d["ms_per_step"], d["value"], d["e2e"]["value"], d["n_gpus"], d.get("final_loss"))
except Exception as e:
    print("ERR", e)
PY
tail -3 gpurun_out/bench_2gpu_ov$m.err
done
