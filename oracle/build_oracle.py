"""Build the C oracle (TEST INFRASTRUCTURE ONLY) -> oracle/_build/libmpc_oracle.so."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libmpc_oracle.so")
SRCS = [os.path.join(HERE, "mpc_oracle.c"), os.path.join(HERE, "mpc_oracle_body.h")]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in SRCS):
        return LIB
    # -march=x86-64-v3 (AVX2+FMA), not -march=native: the .so is built here and travels to the GPU box
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fno-fast-math", "-ffp-contract=fast", "-fopenmp", "-fPIC", "-shared",
           "-o", LIB, SRCS[0], "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
