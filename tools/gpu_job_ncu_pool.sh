#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_k_seg_attention_pool*.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_seg_attention_pool_mma -c 6 -f -o gpurun_out/prof_k_seg_attention_pool_mma python tools/run_c3_tile.py 625000 1 > gpurun_out/ncu_pool.log 2>&1; echo "ncu pool rc=$?"
