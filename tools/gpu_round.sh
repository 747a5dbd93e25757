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
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > $O/smoke.log 2>&1; tail -1 $O/smoke.log
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
# (a window of 50 launches past the pre-pass; tools/make_traffic.py keeps one whole step, k_level0 to k_level0)
timeout 900 ncu --set full --clock-control none -k regex:k_ -s 45 -c 36 -o $O/full_orb \
  python bench.py --steps 2 --warmup 3 --no-ba --no-cpu > $O/ncu_full_orb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 30 -c 12 -o $O/full_ba_local \
  python tools/ba_profile.py local > $O/ncu_full_ba_local.log 2>&1
# nested-dissection kernels of GlobalBA (first LM iteration of the second solve: warm) and of the essential graph
timeout 600 ncu --set full --clock-control none -k regex:k_cr_ -s 46 -c 23 -o $O/full_ba_cr \
  python tools/ba_profile.py global 2 > $O/ncu_full_ba_cr.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches_eg.csv \
  python tools/ba_profile.py essential 1000 > $O/ncu_eg.log 2>&1
python tools/ncu_summary.py $O/full_ba_cr.ncu-rep $O/ncu_full_ba_cr.json > /dev/null 2>&1
python tools/summarize_launches.py $O/launches_eg.csv > $O/launches_eg_summary.txt 2>&1
python -c 'import bench; print(bench.orb_source_hash())' > $O/src_sha256.txt
# summarise on the box: gpurun copies back at most 64 MiB, the reports themselves stay behind when they are large
python tools/ncu_summary.py $O/full_orb.ncu-rep $O/ncu_full_orb.json > /dev/null 2>&1
python tools/ncu_summary.py $O/full_ba_local.ncu-rep $O/ncu_full_ba_local.json > /dev/null 2>&1
for f in launches_orb launches_ba_local launches_ba_global; do python tools/summarize_launches.py $O/$f.csv > $O/${f}_summary.txt 2>&1; done
find $O -name '*.ncu-rep' -size +10M -delete
rm -f $O/launches_eg.csv
ls -la $O
