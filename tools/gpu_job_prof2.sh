#!/bin/bash
# full ncu captures (with source) of the two fine-matching kernels on a 2-tile bench
set -u
mkdir -p gpurun_out
SMALL="python bench.py --tiles 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph --streams 1"
for k in ${NCU_KERNELS:-k_patch_fit_warp k_apply_assign}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k $SMALL > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out
