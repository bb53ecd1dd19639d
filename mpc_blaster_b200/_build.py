"""Build the CUDA library in-tree: mpc_blaster_b200/lib/libmpcb.so (sm_100a only)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libmpcb.so")
SOURCES = ["mpcb_kernels.cu"]
HEADERS = ["mpcb_common.cuh", "mpcb_model.cuh", "mpcb_linearize.cuh", "mpcb_qp.cuh", "mpcb_qp8.cuh", "mpcb_poc.cuh"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "mpcb.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """`defines` / `out` build an experimental variant beside the product library (tools/ab.py)."""
    if out is None and not force and not _stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-warn-spills",
           "-o", out or LIB] + [f"-D{d}" for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
