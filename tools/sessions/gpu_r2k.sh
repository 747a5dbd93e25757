#!/bin/bash
TAG=${1:-r2k}; O=gpurun_out/$TAG; mkdir -p $O
timeout 1200 python -m pytest tests/test_orb_gpu.py tests/test_ref_parity.py tests/test_golden.py tests/test_full_size_gpu.py -x -q -m gpu -k "not global" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -6 $O/pytest.log
for V in 0 1; do CMOS_RESIZE_V1=$V timeout 600 python bench.py --steps 20 --warmup 3 --no-ba --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('RESIZE_V1=$V value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,4) for k,v in d['roofline']['stage_ms'].items()})"; done | tee $O/ab.txt
