"""Build the in-tree CUDA library ``biapy_b200/csrc/libbiapy_b200.so`` for sm_100a with nvcc.

``python -m biapy_b200.build`` (or ``__graft_entry__.build()``).  Objects are rebuilt only when their source
(or a header) is newer; the .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libbiapy_b200.so")
OBJ = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    hdr_m = max(os.path.getmtime(h) for h in hdrs)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            jobs.append((s, o))

    def run(job):
        s, o = job
        r = subprocess.run([nvcc, *NVCC_FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
        with open(o[:-2] + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        return s, r

    failed = False
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(run, jobs):
            if r.returncode != 0:
                failed = True
                sys.stderr.write(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}\n")
            elif verbose:
                sys.stderr.write(r.stderr)
    if failed:
        raise RuntimeError("CUDA build failed")
    if jobs or force or not os.path.exists(LIB):
        r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                            "-Xcompiler", "-fPIC", "-cudart", "static"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
