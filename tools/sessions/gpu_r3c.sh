#!/bin/bash
TAG=${1:-r3c}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_matcher_gpu.py tests/test_full_size_gpu.py tests/test_ba_gpu.py -q -m gpu -x 2>&1 | tail -8 | tee $O/tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-ba --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -c 400 $O/bench.err
python - <<PY
import json
d = json.load(open("$O/bench.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], "arrays", round(d["e2e"]["per_keypoint_arrays"]["value"], 1), d["e2e"]["per_keypoint_arrays"]["h2d_bytes_per_step"])
PY
CMOS_B200_LIB=ceres_mono_orb_slam2_b200/libcmos_b200.so timeout 300 python tools/ba_profile.py global_time 2>&1 | tail -1; timeout 300 python tools/ba_profile.py essential 1000 2>&1 | tail -1
