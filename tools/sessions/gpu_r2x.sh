#!/bin/bash
TAG=${1:-r2x}; O=gpurun_out/$TAG; mkdir -p $O
for LC in "4 16" "6 16" "8 16" "12 16" "8 8" "16 8" "4 32" "6 32" "3 32"; do
  set -- $LC
  timeout 600 python bench.py --steps 20 --warmup 3 --no-ba --no-cpu --split 4 --lanes $1 --chunk $2 > $O/bench_l$1_c$2.json 2> $O/err.txt || tail -3 $O/err.txt
  python - <<PY
import json
d = json.load(open("$O/bench_l$1_c$2.json"))
print("lanes $1 chunk $2: value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["ms_per_step"], "sync", round(d["e2e"]["synchronous"]["value"], 1))
PY
done 2>&1 | tee $O/sweep.txt
timeout 600 python -m pytest tests/test_matcher_gpu.py tests/test_full_size_gpu.py -q -m gpu -x 2>&1 | tail -3
