#!/bin/bash
TAG=${1:-r3g}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_orb_gpu.py tests/test_ref_parity.py tests/test_golden.py tests/test_full_size_gpu.py -q -m gpu -x 2>&1 | tail -5 | tee $O/tests.log
for E in 0; do
CMOS_MATCH_NO_CELLS=$E timeout 600 python bench.py --steps 40 --warmup 3 --no-ba --no-cpu > $O/bench$E.json 2> $O/bench.err; tail -c 300 $O/bench.err
python - <<PY
import json
d = json.load(open("$O/bench$E.json"))
print("no_cells=$E value", round(d["value"], 1), d["ms_per_step"], "single", round(d["single_stream"]["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["form"], {k: round(v, 4) for k, v in d["roofline"]["stage_ms"].items()})
PY
done
