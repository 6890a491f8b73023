#!/bin/bash
# L2 prefetch warp / resident-weight mode of the x-slab kernels: parity tests + A/B micro timings + bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -q -x > gpurun_out/pytest_umma.log 2>&1; echo "umma rc=$?"; tail -4 gpurun_out/pytest_umma.log
for pf in 0 1 2 4; do
  B200_XSLAB_PF=$pf timeout 300 python tools/epi_micro.py > gpurun_out/pf_micro_$pf.log 2>&1
  grep "EPI_TMA" gpurun_out/pf_micro_$pf.log | sed "s/^/[PF=$pf] /"
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err; echo "bench rc=$?"
B200_XSLAB_PF=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pf0.json 2> gpurun_out/bench_pf0.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench_pf", "bench_pf0"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["final_loss"], d["roofline"]["frac"], d["roofline"]["achieved"])
    except Exception as e:
        print(f, "ERR", e)
PY
