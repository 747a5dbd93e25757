#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch lists and --set full captures. Outputs under gpurun_out/.
# usage: tools/gpu_round.sh <tag> [tests|notests]
TAG=${1:-x}; MODE=${2:-tests}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
if [ "$MODE" = tests ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
  tail -5 $O/pytest.log
fi
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
tail -c 600 $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>> $O/bench.err
# launch lists (shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $O/launches_orb.csv \
  python bench.py --steps 2 --warmup 3 --no-ba --no-cpu > $O/ncu_orb.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_ba_local.csv \
  python tools/ba_profile.py local > $O/ncu_ba_local.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_ba_global.csv \
  python tools/ba_profile.py global 2 > $O/ncu_ba_global.log 2>&1
# full captures: one step of the ORB+match path, every kernel once
# (pre-pass extraction 12 launches + 3 warm-up steps x 15 = 57 launches before the first timed 64-frame step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 57 -c 15 -o $O/full_orb \
  python bench.py --steps 2 --warmup 3 --no-ba --no-cpu > $O/ncu_full_orb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 30 -c 12 -o $O/full_ba_local \
  python tools/ba_profile.py local > $O/ncu_full_ba_local.log 2>&1
ls -la $O
