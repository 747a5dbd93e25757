#!/bin/bash
TAG=${1:-r2y}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_ba_gpu.py -q -m gpu -k "essential or loop_closing" 2>&1 | tail -30 | tee $O/eg_tests.log
timeout 300 python tools/ba_profile.py essential 1000 2>&1 | tail -3 | tee $O/eg_time.txt
CMOS_EG_BLOCKED=1 timeout 300 python tools/ba_profile.py essential 1000 2>&1 | tail -2 | tee -a $O/eg_time.txt
