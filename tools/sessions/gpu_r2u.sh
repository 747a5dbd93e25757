#!/bin/bash
TAG=${1:-r2u}; O=gpurun_out/$TAG; mkdir -p $O
for L in "$@"; do :; done
for LIB in build/libcmos_choltiming.so; do
CMOS_B200_LIB=$LIB python - <<'PY' 2>&1 | tee -a $O/timing.txt
import ctypes as C, numpy as np
from ceres_mono_orb_slam2_b200 import _lib
L = _lib.lib()
for n in (120, 144):
    rng = np.random.default_rng(n); B = rng.standard_normal((n, n)); A = B @ B.T / n + np.eye(n); b = np.ones(n)
    x = np.zeros(n); f = C.c_int32(); cyc = (C.c_int64 * 2)()
    _lib.check(L.cmos_debug_solve_spd(_lib.ptr(A), _lib.ptr(b), n, 0, _lib.ptr(x), C.byref(f), cyc))
    print(f"n={n}: factor+invert24 {cyc[0]} cycles, back substitution {cyc[1]} cycles, failed {f.value}, err {np.abs(x - np.linalg.solve(A, b)).max():.2e}")
PY
done
for S in 1 2 4; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-ba --no-cpu --split $S > $O/bench_split$S.json 2> $O/bench_split$S.err; echo "split $S exit $?"
  python - <<PY
import json
d = json.load(open("$O/bench_split$S.json"))
print("split $S:", d["value"], d["ms_per_step"], "single", d["single_stream"]["value"], "e2e", d["e2e"]["value"], d["clocks"]["sm_mhz"])
PY
done
