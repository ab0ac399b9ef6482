#!/bin/bash
set -u
mkdir -p gpurun_out
for k in k_kabsch_moments k_apply_transforms k_grid_search k_rigidity; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
