#!/bin/bash
# BA-focused GPU pass: parity tests of the optimiser, bench line, LocalBA launch list.
O=gpurun_out/${1:-ba2}; mkdir -p $O
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_golden.py tests/test_cpp_adapters.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
tail -c 300 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
for k,v in d["ba"].items():
    if isinstance(v, dict) and "value" in v: print(k, v["value"], v.get("ms_per_solve"), v.get("gpu_launches_per_solve"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_ba_local.csv \
  python tools/ba_profile.py local > $O/ncu_ba_local.log 2>&1
python tools/summarize_launches.py $O/launches_ba_local.csv | head -16
