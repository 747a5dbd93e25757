#!/bin/bash
TAG=${1:-r3d}; O=gpurun_out/$TAG; mkdir -p $O
for i in 1 2 3; do
timeout 600 python bench.py --steps 40 --warmup 3 --no-ba --no-cpu > $O/bench$i.json 2> $O/bench.err
python - <<PY
import json
d = json.load(open("$O/bench$i.json"))
print("run $i: value", round(d["value"], 1), "e2e compact", round(d["e2e"]["value"], 1), d["e2e"]["ms_per_step"], "arrays", round(d["e2e"]["per_keypoint_arrays"]["value"], 1))
PY
done
