#!/bin/bash
TAG=${1:-r2g}; O=gpurun_out/$TAG; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -6 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -c 400 $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>> $O/bench.err; echo "ref exit $?"
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"cpu",d.get("cpu_baseline"))
print("roofline",{k:v for k,v in d["roofline"].items() if k in("kernel","frac","traffic","traffic_source","largest_single_kernel","stage_ms")})
print("configs0",d.get("configs0"))
print("secondary",d.get("secondary"))
for k,v in d["ba"].items():
    if isinstance(v,dict) and "value" in v: print(k, round(v["value"],1), round(v.get("ms_per_solve",0),3))
r=json.load(open("$O/bench_ref.json")); print("ref", r["value"], r["cpu_baseline"])
PY
