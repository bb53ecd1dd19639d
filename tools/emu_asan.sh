#!/bin/sh
# compute-sanitizer is closed on this GPU pool, so memory-safety of the kernel bodies is checked on
# the host: the same .cuh sources compiled with -DMPCB_HOST_EMU under AddressSanitizer + UBSan
# (workspace, per-warp shared-memory struct and register arrays are all bounds-checked).
set -e
cd "$(dirname "$0")/.."
g++ -O1 -g -std=c++17 -DMPCB_HOST_EMU -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -shared \
    -Itests/emu -Impc_blaster_b200/csrc -x c++ tests/emu/emu_main.cpp -o /tmp/libmpcb_emu_asan.so
cat > /tmp/mpcb_asan_run.py <<'PY'
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/emu')
import emu_binding as eb
eb._lib = C.CDLL('/tmp/libmpcb_emu_asan.so'); eb._lib.emu_params_size.restype = C.c_size_t
from oracle import blaster_oracle as bo
from mpc_blaster_b200 import scenarios as sc
for var in (17, 12):
    P = bo.canonical_problem(5, var)
    x0, yref = sc.random_setpoints(1, seed=3, nx=P.nx, nu=P.nu)
    X = np.repeat(x0, 6, axis=0).copy(); U = np.tile(sc.hover_trim(P.nu), (5, 1)).copy()
    print(var, eb.rti_solve(P, X, U, x0[0], yref[0], bo.default_params())[:2])
    # four instances per warp (mpcb_qp8.cuh): a full warp, a partly filled one, and one whose groups are refilled
    for nb in (4, 3, 9):
        x0, yref = sc.random_setpoints(nb, seed=4, nx=P.nx, nu=P.nu)
        X = np.repeat(x0[:, None, :], 6, axis=1).copy(); U = np.tile(sc.hover_trim(P.nu), (nb, 5, 1)).copy()
        print(var, 'qp8', nb, eb.rti_solve4(P, X, U, x0, yref, bo.default_params()))
print('poc', eb.poc([0, -0.05, 0], [0.2117, 0], [0.6, 0, 3.5], mode=0)[0], eb.poc([0, -0.05, 0], [0.2117, 0], [0.6, 0, 3.5], mode=1)[0])
PY
LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" \
    ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python /tmp/mpcb_asan_run.py
