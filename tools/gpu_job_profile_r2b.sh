#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
BARGS="--tiles 8 --steps 2 --warmup 1 --configs none --no-cpu-baseline --no-e2e --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(k_|.*cub).*' -c 900 --csv --log-file gpurun_out/launches.csv python bench.py $BARGS > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_desc_nn_tc -s 2 -c 1 -f -o gpurun_out/prof_k_desc_nn_tc python tools/bench_desc.py --n 151552 --m 262144 --d 64 --reps 1 > gpurun_out/ncu_desc.log 2>&1; echo "ncu desc64 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_desc_nn_tc -s 2 -c 1 -f -o gpurun_out/prof_k_desc_nn_tc_d32 python tools/bench_desc.py --n 151552 --m 262144 --d 32 --reps 1 > gpurun_out/ncu_desc32.log 2>&1; echo "ncu desc32 rc=$?"
timeout 600 python -m pytest tests/test_dips_gpu.py tests/test_host_api_gpu.py tests/test_knn_gpu.py tests/test_entry_gpu.py -m gpu -q -x 2>&1 | tail -5
