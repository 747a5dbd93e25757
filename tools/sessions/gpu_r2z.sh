#!/bin/bash
TAG=${1:-r2z}; O=gpurun_out/$TAG; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_ba_local.csv python tools/ba_profile.py local > $O/ncu_ba_local.log 2>&1
python tools/summarize_launches.py $O/launches_ba_local.csv | head -14 | tee $O/launches_ba_local_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_ba_global.csv python tools/ba_profile.py global 2 > $O/ncu_ba_global.log 2>&1
python tools/summarize_launches.py $O/launches_ba_global.csv | head -22 | tee $O/launches_ba_global_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches_eg.csv python tools/ba_profile.py essential 1000 > $O/ncu_eg.log 2>&1
python tools/summarize_launches.py $O/launches_eg.csv | head -24 | tee $O/launches_eg_summary.txt
rm -f $O/*.csv
timeout 600 python -m pytest tests/test_ba_gpu.py -q -m gpu -k "nested_dissection" 2>&1 | tail -3
