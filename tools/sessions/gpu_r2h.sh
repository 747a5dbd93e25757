#!/bin/bash
TAG=${1:-r2h}; O=gpurun_out/$TAG; mkdir -p $O
timeout 1500 python -m pytest tests/test_matcher_gpu.py tests/test_ref_matcher_parity.py tests/test_ref_parity.py tests/test_cpp_adapters.py -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -6 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-ba > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -c 300 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"], d["e2e"]["ms_per_step"], "sync", d["e2e"]["synchronous"])
PY
for L in "2 32" "4 16" "8 8" "4 8" "6 12"; do set -- $L; python bench.py --steps 20 --warmup 3 --no-ba --no-cpu --lanes $1 --chunk $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lanes $1 chunk $2: e2e', round(d['e2e']['value'],1), 'sync', round(d['e2e']['synchronous']['value'],1))"; done | tee $O/e2e_sweep.txt
