#!/bin/bash
# A/B of the x-slab epilogues (direct stores vs TMA store, fused channel statistics): parity tests, micro timings, bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -q > gpurun_out/pytest_umma.log 2>&1; echo "umma rc=$?"; tail -8 gpurun_out/pytest_umma.log
B200_EPI_TMA=1 timeout 300 python tools/epi_micro.py > gpurun_out/epi_micro.log 2>&1
B200_EPI_TMA=1 B200_EPI_SWZ=64 timeout 300 python tools/epi_micro.py >> gpurun_out/epi_micro.log 2>&1
B200_EPI_TMA=1 B200_EPI_SWZ=0 timeout 300 python tools/epi_micro.py >> gpurun_out/epi_micro.log 2>&1
grep -v "^ \|Traceback\|torch\.\|CUDA\|^$\|^Search\|^For\|^Compile" gpurun_out/epi_micro.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_umma.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --detail > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
B200_EPI_TMA=0 B200_FUSE_STATS=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_legacy.json 2> gpurun_out/bench_legacy.err; echo "bench legacy rc=$?"
B200_EPI_TMA=1 B200_FUSE_STATS=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_tma_nofuse.json 2> gpurun_out/bench_tma_nofuse.err; echo "bench tma-nofuse rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_legacy", "bench_tma_nofuse"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["final_loss"])
    except Exception as e:
        print(f, "ERR", e)
PY
