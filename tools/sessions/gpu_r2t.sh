#!/bin/bash
# new register-resident Cholesky: solver test alone first (all cases), A/B against the shared-memory build, then the round
TAG=${1:-r2t}; O=gpurun_out/$TAG; mkdir -p $O
timeout 600 python -m pytest tests/test_solver_gpu.py -q -m gpu 2>&1 | tail -25 | tee $O/solver_test.log
python - <<'PY' 2>&1 | tee $O/solver_cycles.txt
import ctypes as C, numpy as np
from ceres_mono_orb_slam2_b200 import _lib
L = _lib.lib()
for n in (24, 60, 114, 120, 144, 150, 156, 228):
    rng = np.random.default_rng(n); B = rng.standard_normal((n, n)); A = B @ B.T / n + np.eye(n); b = np.ones(n)
    x = np.zeros(n); f = C.c_int32(); cyc = (C.c_int64 * 2)()
    _lib.check(L.cmos_debug_solve_spd(_lib.ptr(A), _lib.ptr(b), n, 0, _lib.ptr(x), C.byref(f), cyc))
    print(f"n={n}: factor+invert24 {cyc[0]} cycles, back substitution {cyc[1]} cycles, failed {f.value}, err {np.abs(x - np.linalg.solve(A, b)).max():.2e}")
PY
for L in ceres_mono_orb_slam2_b200/libcmos_b200.so build/libcmos_cholsmem.so; do
  timeout 600 python tools/ab_solve.py $L 2>&1 | tail -1
  echo $L; CMOS_B200_LIB=$L timeout 600 python tools/ba_profile.py global_time 2>&1 | tail -1
done 2>&1 | tee $O/ab.txt
