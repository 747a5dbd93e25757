#!/bin/bash
TAG=${1:-r2n}; O=gpurun_out/$TAG; mkdir -p $O
timeout 1200 python -m pytest tests/test_orb_gpu.py tests/test_ref_parity.py tests/test_golden.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-ba --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,4) for k,v in d['roofline']['stage_ms'].items()})" | tee $O/bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 16 --csv --log-file $O/launches_orb.csv python bench.py --steps 2 --warmup 3 --no-ba --no-cpu > $O/ncu.log 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_orb.csv")) if len(r)>5]
start=next(i for i,r in enumerate(rows) if r[0]=="ID"); hdr=rows[start]
for r in rows[start+1:start+20]:
    rec=dict(zip(hdr,r))
    if rec.get("Metric Name")=="gpu__time_duration.sum": print(rec["Kernel Name"][:36], rec["Grid Size"], rec["Block Size"], rec["Metric Value"], rec["Metric Unit"])
PY
