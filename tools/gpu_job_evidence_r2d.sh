#!/bin/bash
# final evidence of round 2: launch list of one C5 step (final code), ncu --set full of the 8-warp pooling kernel
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_k_seg_attention_pool*.ncu-rep gpurun_out/launches.csv
BARGS="--tiles 8 --steps 2 --warmup 1 --configs none --no-cpu-baseline --no-e2e --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(k_|.*cub).*' -c 900 --csv --log-file gpurun_out/launches.csv python bench.py $BARGS > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_seg_attention_pool_mma -c 6 -f -o gpurun_out/prof_k_seg_attention_pool_mma python tools/run_c3_tile.py 625000 1 > gpurun_out/ncu_pool.log 2>&1; echo "ncu pool rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_patch_fit_warp\$" -s 10 -c 1 -f -o gpurun_out/prof_k_patch_fit_warp python bench.py --tiles 4 --steps 1 --warmup 1 --configs none --no-cpu-baseline --no-e2e --no-graph --streams 1 > gpurun_out/ncu_fit.log 2>&1; echo "ncu fit rc=$?"
