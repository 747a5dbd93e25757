#!/bin/bash
TAG=${1:-r2i}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_full_size_gpu.py -x -q -k "not extract" > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
for L in ceres_mono_orb_slam2_b200/libcmos_b200.so build/libcmos_cholv1.so; do
  echo $L; CMOS_B200_LIB=$L timeout 600 python tools/ba_profile.py global_time 2>&1 | tail -1
  timeout 600 python tools/ab_solve.py $L 2>&1 | tail -1
done 2>&1 | tee $O/ab.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file $O/launches_local_warm.csv python tools/ba_profile.py local > $O/ncu_local.log 2>&1
python tools/summarize_launches.py $O/launches_local_warm.csv > $O/launches_local_warm_summary.txt 2>&1; head -8 $O/launches_local_warm_summary.txt
