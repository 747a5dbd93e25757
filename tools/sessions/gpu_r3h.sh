#!/bin/bash
TAG=${1:-r3h}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_resize2|k_blur|k_sf_replay|k_octree|k_level0|k_borders|k_build_grid" -s 13 -c 13 -o $O/src_orb_rest \
  python bench.py --steps 2 --warmup 3 --no-ba --no-cpu --split 1 > $O/ncu_src.log 2>&1
ls -la $O/*.ncu-rep
ncu -i $O/src_orb_rest.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; c={n:i for i,n in enumerate(h)}
for r in rows[2:]:
    print(r[c['Kernel Name']][:30], 'inst', r[c['smsp__inst_executed.sum']], 'dur_us', r[c['gpu__time_duration.sum']], 'issue%', r[c['sm__issue_active.avg.pct_of_peak_sustained_elapsed']][:5], 'grid', r[c['launch__grid_size']], 'regs', r[c['launch__registers_per_thread']], 'dram%', r[c['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']][:5], 'occ%', r[c['sm__warps_active.avg.pct_of_peak_sustained_active']][:5])
"
