#!/bin/bash
# 8-GPU visit: the bench line at N = 8 (training weak scaling + ONE 512^3 volume sharded over 8 ranks) and at N = 4
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"
tail -5 gpurun_out/bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_{n}gpu.json") if l.startswith("{")][-1])
        i = d["infer"]
        print(n, "GPUs train:", round(d["ms_per_step"], 3), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
              "| infer ms", round(i["ms_per_volume"], 2), round(i["value"], 1), "e2e", round(i["e2e"]["value"], 1), i["config"]["patches_this_rank"],
              i["config"]["exchange_bytes_received_rank0"])
    except Exception as e:
        print(n, "ERR", e)
PY
