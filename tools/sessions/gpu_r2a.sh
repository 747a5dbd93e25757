#!/bin/bash
# round 2, pass A: reference-parity tests, BA tests with the nested-dissection solver, full-size tests, BA bench numbers
TAG=${1:-r2a}; O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version --format=csv > $O/gpu.txt
timeout 600 python -m pytest tests/test_ref_parity.py -x -q -m gpu > $O/pytest_ref.log 2>&1; echo "exit $?" >> $O/pytest_ref.log; tail -4 $O/pytest_ref.log
timeout 900 python -m pytest tests/test_ba_gpu.py -x -q > $O/pytest_ba.log 2>&1; echo "exit $?" >> $O/pytest_ba.log; tail -6 $O/pytest_ba.log
timeout 1200 python -m pytest tests/test_full_size_gpu.py -x -q -k "global" > $O/pytest_full.log 2>&1; echo "exit $?" >> $O/pytest_full.log; tail -12 $O/pytest_full.log
timeout 600 python tools/ba_profile.py global_time > $O/global_time.txt 2>&1; tail -3 $O/global_time.txt
CMOS_BA_BAND_SERIAL=1 timeout 600 python tools/ba_profile.py global_time > $O/global_time_serial.txt 2>&1; tail -2 $O/global_time_serial.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_global.csv python tools/ba_profile.py global 2 > $O/ncu_global.log 2>&1
python tools/summarize_launches.py $O/launches_global.csv > $O/launches_global_summary.txt 2>&1; head -40 $O/launches_global_summary.txt
