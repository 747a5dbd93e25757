#!/bin/bash
# ncu --set full of the global-BA kernels (one launch each, second LM iteration), summaries to gpurun_out/<tag>/
TAG=${1:-ncu_ba}; O=gpurun_out/$TAG; mkdir -p $O
for K in k_schur_warp k_cr_factor k_cr_spike k_cr_back k_cr_schur k_cr_assemble k_backsub k_cam_blocks k_linearize; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K\$ --launch-skip 1 --launch-count 1 -f -o $O/full_$K python tools/ba_profile.py global 2 > $O/ncu_$K.log 2>&1
done
for K in k_schur_warp k_cr_factor k_cr_spike k_cr_back k_cr_schur k_cr_assemble k_backsub k_cam_blocks k_linearize; do
  python tools/ncu_summary.py $O/full_$K.ncu-rep $O/full_$K.json | tail -1
done
