#!/bin/bash
TAG=${1:-r2q}; O=gpurun_out/$TAG; mkdir -p $O
for L in ceres_mono_orb_slam2_b200/libcmos_b200.so build/libcmos_cam256.so build/libcmos_cam512.so; do
  timeout 600 python tools/ab_solve.py $L 2>&1 | tail -1
  echo $L; CMOS_B200_LIB=$L timeout 600 python tools/ba_profile.py global_time 2>&1 | tail -1
done 2>&1 | tee $O/ab.txt
CMOS_B200_LIB=build/libcmos_crtiming.so timeout 300 python tools/ba_profile.py local 2>&1 | grep -E "solve_small|packed_chol" | head -6
