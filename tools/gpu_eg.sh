#!/bin/bash
# GPU pass for the loop-closing solves: parity tests, timing, launch list.
O=gpurun_out/${1:-eg}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python tools/ba_profile.py essential 1000 > $O/essential.log 2>&1; tail -3 $O/essential.log
timeout 300 python tools/ba_profile.py essential 200 >> $O/essential.log 2>&1; tail -2 $O/essential.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_essential.csv \
  python tools/ba_profile.py essential 1000 > $O/ncu_essential.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file $O/launches_sim3.csv \
  python tools/ba_profile.py sim3 > $O/ncu_sim3.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -3 $O/smoke.log
