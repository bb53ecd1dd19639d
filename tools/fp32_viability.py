"""Experiment (test infrastructure, not product): can the QP of this path be solved in pure FP32?

north_star lists an FP32 tolerance (|du|,|dx| <= 1e-4) and config 5 an FP32 column.  Before writing an
FP32 instantiation of the kernels, this script answers the numerical question on the CPU: it rebuilds the
C oracle (oracle/mpc_oracle.c, the same Riccati-LQ Mehrotra iteration the CUDA kernels run) with every
`double` replaced by `float`, solves configs[1]'s first 256 instances, and compares with the FP64 oracle
at several KKT tolerances.  Output recorded in profiles/r01_fp32_viability.txt; conclusion in DESIGN.md
section 8.  Build products go to gpurun_out/fp32/ (scratch, git-ignored).

    python tools/fp32_viability.py
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpc_blaster_b200 import scenarios as sc  # noqa: E402
from oracle import blaster_oracle as bo, c_oracle as co  # noqa: E402

F = C.c_float
OUT = os.path.join(ROOT, "gpurun_out", "fp32")


def build32() -> C.CDLL:
    os.makedirs(OUT, exist_ok=True)
    for name, dst in (("mpc_oracle.c", "orc32.c"), ("mpc_oracle_body.h", "mpc_oracle_body.h")):
        src = open(os.path.join(ROOT, "oracle", name)).read()
        open(os.path.join(OUT, dst), "w").write(re.sub(r"\bdouble\b", "float", src))
    lib = os.path.join(OUT, "liborc32.so")
    subprocess.check_call(["gcc", "-O2", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-o", lib,
                           os.path.join(OUT, "orc32.c"), "-lm"])
    return C.CDLL(lib)


class P32(C.Structure):
    _fields_ = [(n, {C.c_double: F}.get(t, t) if not hasattr(t, "_length_") else F * t._length_)
                for n, t in co.OrcProblem._fields_]


def to32(o64: co.OrcProblem) -> P32:
    o = P32()
    for name, _ in P32._fields_:
        v = getattr(o64, name)
        if hasattr(v, "__len__"):
            a = getattr(o, name)
            for i in range(len(v)):
                a[i] = v[i]
        else:
            setattr(o, name, v)
    return o


def fp(a):
    return a.ctypes.data_as(C.POINTER(F))


def solve32(lib, P, x0, yref, trim, **kw):
    B = x0.shape[0]
    o = to32(co.make_problem(P, **kw))
    X = np.zeros((B, P.N + 1, P.nx), np.float32)
    X[:] = x0.reshape(B, 1, -1)
    U = np.zeros((B, P.N, P.nu), np.float32)
    U[:] = trim.reshape(1, 1, -1)
    st = np.zeros(B, np.int32)
    it = np.zeros(B, np.int32)
    rc = lib.orc_rti_solve_batch(C.byref(o), fp(X), fp(U), fp(x0.astype(np.float32)), fp(yref.astype(np.float32)), 1,
                                 fp(bo.default_params().astype(np.float32)), 0,
                                 st.ctypes.data_as(C.POINTER(C.c_int32)), it.ctypes.data_as(C.POINTER(C.c_int32)), B, 8)
    assert rc == 0
    return X, U, st, it


def main():
    lib = build32()
    assert lib.orc_problem_size() == C.sizeof(P32)
    B, N = 256, 20
    x0, yref = sc.random_setpoints(B, seed=1234)
    trim = sc.hover_trim()
    P = bo.canonical_problem(N)
    orc = co.BatchRTI(P, B)
    orc.reset(x0, trim)
    _, X64, U64, st64 = orc.solve(x0, yref)
    print(f"FP64 oracle, HPIPM-default tolerances: converged {np.mean(st64 == 0):.3f}, IPM iterations mean "
          f"{orc.iters.mean():.2f} max {orc.iters.max()}")
    print("FP32 build of the same iteration (status 3 = min-step, 4 = factorisation breakdown):")
    for tol in ((1e-6, 1e-8, 1e-8, 1e-8), (1e-3, 1e-5, 1e-5, 1e-5), (1e-2, 1e-4, 1e-4, 1e-6), (1e-2, 1e-4, 1e-4, 1e-4),
                (1e-1, 1e-4, 1e-4, 1e-3)):
        kw = dict(zip(("tol_stat", "tol_eq", "tol_ineq", "tol_comp"), tol))
        X, U, st, it = solve32(lib, P, x0, yref, trim, **kw)
        ok = (st == 0) & (st64 == 0)
        dU, dX = np.abs(U[ok] - U64[ok]), np.abs(X[ok] - X64[ok])
        print(f"  tol stat/eq/ineq/comp = {tol}: converged {np.mean(st == 0):.3f} status histogram "
              f"{np.bincount(st, minlength=5).tolist()} iterations mean {it.mean():.1f} max {it.max()}")
        if ok.any():
            print(f"      vs FP64 on the converged ones: max|dU| thrust {dU[..., :4].max():.2e} swivel {dU[..., 4:].max():.2e}, "
                  f"u0 thrust {dU[:, 0, :4].max():.2e}, max|dX| {dX.max():.2e}, median over instances of max|dU| "
                  f"{np.median(dU.reshape(int(ok.sum()), -1).max(1)):.2e}   (north_star's FP32 bar: 1e-4)")


if __name__ == "__main__":
    main()
