#!/bin/bash
TAG=${1:-r2f}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_full_size_gpu.py -x -q -k "not extract" > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
timeout 600 python tools/ba_profile.py global_time 2>&1 | tail -1 | tee $O/global_time.txt
timeout 600 python tools/ab_solve.py ceres_mono_orb_slam2_b200/libcmos_b200.so 2>&1 | tail -1 | tee $O/local_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/launches_global_warm.csv python tools/ba_profile.py global 2 > $O/ncu_global.log 2>&1
python tools/summarize_launches.py $O/launches_global_warm.csv > $O/launches_global_warm_summary.txt 2>&1; head -12 $O/launches_global_warm_summary.txt
