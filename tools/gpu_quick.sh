#!/bin/bash
# quick GPU pass: ORB + matcher parity tests, bench (no BA), optional ncu of selected kernels.
# usage: tools/gpu_quick.sh <tag> [kernel-regex for ncu --set full]
TAG=${1:-q}; KRE=${2:-}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_orb_gpu.py tests/test_matcher_gpu.py -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -15 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-ba --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
tail -c 1500 $O/bench.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench.json")); print(round(d["value"],2), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, "e2e", round(d["e2e"]["value"],2), d["e2e"]["ms_per_step"])
except Exception as e: print("failed", e)
PY
if [ -n "$KRE" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 60 -c 8 -o $O/full \
    python bench.py --steps 1 --warmup 3 --no-ba --no-cpu > $O/ncu_full.log 2>&1
  tail -3 $O/ncu_full.log
fi
