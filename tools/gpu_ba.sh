#!/bin/bash
# BA pass: parity tests + BA bench numbers. usage: tools/gpu_ba.sh <tag>
TAG=${1:-ba}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_ba_gpu.py -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -8 $O/pytest.log
timeout 900 python - > $O/ba.json 2> $O/ba.err <<PY
import json, argparse, sys
sys.path.insert(0, '.')
from ceres_mono_orb_slam2_b200 import ba_bench
a = argparse.Namespace(steps=10, no_global=False, global_iters=10)
print(json.dumps(ba_bench.run(0, 1, a)))
PY
tail -c 800 $O/ba.err
python - <<PY
import json
d=json.load(open("$O/ba.json"))
for k,v in d.items():
    if isinstance(v,dict) and 'value' in v: print(k, round(v['value'],1), 'Mresid/s', round(v['ms_per_solve'],3), 'ms', v.get('iterations'), v.get('cost',[[0,0]])[-1], 'launches', v.get('gpu_launches_per_solve'))
PY
