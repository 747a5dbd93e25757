#!/bin/bash
TAG=${1:-r3a}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_full_size_gpu.py tests/test_cpp_adapters.py -q -m gpu -x 2>&1 | tail -8 | tee $O/tests.log
timeout 900 python - <<'PY' 2>&1 | tee $O/ba_bench.txt
import json, argparse
from ceres_mono_orb_slam2_b200 import ba_bench
a = argparse.Namespace(steps=20, warmup=3, no_global=False, global_iters=10)
r = ba_bench.run(0, 1, a)
for k, v in r.items():
    if isinstance(v, dict) and "value" in v:
        print(k, round(v["value"], 1), v.get("ms_per_solve"), "e2e", v.get("e2e"))
    else:
        print(k, str(v)[:300])
PY
