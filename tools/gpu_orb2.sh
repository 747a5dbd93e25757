#!/bin/bash
# ORB-focused GPU pass: extractor / matcher parity tests, bench with and without the blur/quadtree overlap.
O=gpurun_out/${1:-orb2}; mkdir -p $O
timeout 900 python -m pytest tests/test_orb_gpu.py tests/test_matcher_gpu.py tests/test_golden.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-ba --no-cpu > $O/bench_overlap.json 2> $O/bench.err; echo "bench exit $?"
CMOS_ORB_NO_OVERLAP=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-ba --no-cpu > $O/bench_serial.json 2>> $O/bench.err
python - <<PY
import json
for n in ("overlap", "serial"):
    d = json.load(open("$O/bench_%s.json" % n))
    print(n, round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()})
PY
