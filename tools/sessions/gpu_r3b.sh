#!/bin/bash
TAG=${1:-r3b}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_ba_gpu.py tests/test_full_size_gpu.py tests/test_cpp_adapters.py -q -m gpu 2>&1 | tail -8 | tee $O/tests.log
python - <<'PY' 2>&1 | tee $O/cycles.txt
import ctypes as C, numpy as np
from ceres_mono_orb_slam2_b200 import _lib
L = _lib.lib()
for n in (114, 120):
  for variant in (0, 1):
    rng = np.random.default_rng(n); B = rng.standard_normal((n, n)); A = B @ B.T / n + np.eye(n); b = np.ones(n)
    x = np.zeros(n); f = C.c_int32(); cyc = (C.c_int64 * 2)()
    _lib.check(L.cmos_debug_solve_spd(_lib.ptr(A), _lib.ptr(b), n, variant, _lib.ptr(x), C.byref(f), cyc))
    print(f"n={n} variant {variant}: {cyc[0]} + {cyc[1]} cycles, failed {f.value}, err {np.abs(x - np.linalg.solve(A, b)).max():.2e}")
PY
timeout 900 python - <<'PY' 2>&1 | tee $O/ba_bench.txt
import json, argparse
from ceres_mono_orb_slam2_b200 import ba_bench
a = argparse.Namespace(steps=20, warmup=3, no_global=True, global_iters=10)
r = ba_bench.run(0, 1, a)
for k, v in r.items():
    if isinstance(v, dict) and "value" in v:
        print(k, round(v["value"], 1), v.get("ms_per_solve"), "e2e", v.get("e2e"), v.get("gpu_launches_per_solve"))
PY
