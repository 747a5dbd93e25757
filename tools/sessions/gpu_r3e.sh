#!/bin/bash
TAG=${1:-r3e}; O=gpurun_out/$TAG; mkdir -p $O
for S in 8 4; do
timeout 600 python bench.py --steps 40 --warmup 3 --no-ba --no-cpu --split $S > $O/bench_s$S.json 2> $O/bench.err
python - <<PY
import json
d = json.load(open("$O/bench_s$S.json"))
print("split $S: value", round(d["value"], 1), d["ms_per_step"], "e2e", round(d["e2e"]["value"], 1), d["e2e"]["form"])
PY
done
# rare paths (keyframe searches, vocabulary transform, map-point maintenance, PoseOptimization, OptimizeSim3): launch list
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_rare.csv \
  python -m pytest tests/test_matcher_kf_gpu.py -q -m gpu -x > $O/ncu_rare.log 2>&1
python tools/summarize_launches.py $O/launches_rare.csv > $O/launches_rare_kf_summary.txt 2>&1; head -30 $O/launches_rare_kf_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_rare2.csv \
  python -c "
import sys; sys.argv=['x','pose']
exec(open('tools/ba_profile.py').read())
" > $O/ncu_rare2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_rare3.csv python tools/ba_profile.py sim3 > $O/ncu_rare3.log 2>&1
python tools/summarize_launches.py $O/launches_rare2.csv > $O/launches_pose_summary.txt 2>&1; head -5 $O/launches_pose_summary.txt
python tools/summarize_launches.py $O/launches_rare3.csv > $O/launches_sim3_summary.txt 2>&1; head -5 $O/launches_sim3_summary.txt
rm -f $O/*.csv
# source-level counters of the three heaviest ORB / matcher kernels (one launch each, 64 frames)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_describe|k_sf_lists|k_fast" -s 9 -c 3 -o $O/src_orb3 \
  python bench.py --steps 2 --warmup 3 --no-ba --no-cpu --split 1 > $O/ncu_src.log 2>&1
ls -la $O/*.ncu-rep
