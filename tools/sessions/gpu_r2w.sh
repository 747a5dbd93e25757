#!/bin/bash
TAG=${1:-r2w}; O=gpurun_out/$TAG; mkdir -p $O
CMOS_B200_LIB=build/libcmos_choltiming.so python - <<'PY' 2>&1 | tee $O/timing.txt
import ctypes as C, numpy as np
from ceres_mono_orb_slam2_b200 import _lib
L = _lib.lib()
for n in (120,):
    rng = np.random.default_rng(n); B = rng.standard_normal((n, n)); A = B @ B.T / n + np.eye(n); b = np.ones(n)
    x = np.zeros(n); f = C.c_int32(); cyc = (C.c_int64 * 2)()
    _lib.check(L.cmos_debug_solve_spd(_lib.ptr(A), _lib.ptr(b), n, 0, _lib.ptr(x), C.byref(f), cyc))
    print(f"n={n}: factor+invert24 {cyc[0]} cycles, back substitution {cyc[1]} cycles, failed {f.value}, err {np.abs(x - np.linalg.solve(A, b)).max():.2e}")
PY
