#!/bin/bash
TAG=${1:-r2d}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_full_size_gpu.py -x -q -k "not extract" > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
timeout 600 python tools/ba_profile.py global_time > $O/global_time.txt 2>&1; tail -1 $O/global_time.txt
CMOS_BA_SCHUR=cta timeout 600 python tools/ba_profile.py global_time > $O/global_time_schur_cta.txt 2>&1; tail -1 $O/global_time_schur_cta.txt
timeout 600 python tools/ab_solve.py ceres_mono_orb_slam2_b200/libcmos_b200.so > $O/local_time.txt 2>&1; tail -1 $O/local_time.txt
CMOS_BA_SCHUR=warp timeout 600 python tools/ab_solve.py ceres_mono_orb_slam2_b200/libcmos_b200.so > $O/local_time_schur_warp.txt 2>&1; tail -1 $O/local_time_schur_warp.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/launches_global_warm.csv python tools/ba_profile.py global 2 > $O/ncu_global.log 2>&1
python tools/summarize_launches.py $O/launches_global_warm.csv > $O/launches_global_warm_summary.txt 2>&1; head -20 $O/launches_global_warm_summary.txt
