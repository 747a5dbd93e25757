#!/bin/bash
TAG=${1:-r2b}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_full_size_gpu.py -x -q -k "global" > $O/pytest_full.log 2>&1; echo "exit $?" >> $O/pytest_full.log; tail -5 $O/pytest_full.log
timeout 600 python tools/ba_profile.py global_time > $O/global_time.txt 2>&1; tail -3 $O/global_time.txt
CMOS_B200_LIB=build/libcmos_crtiming.so timeout 600 python tools/ba_profile.py global 2 > $O/crtiming.txt 2>&1; grep k_cr_factor $O/crtiming.txt | head -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_global.csv python tools/ba_profile.py global 2 > $O/ncu_global.log 2>&1
python tools/summarize_launches.py $O/launches_global.csv > $O/launches_global_summary.txt 2>&1; head -20 $O/launches_global_summary.txt
