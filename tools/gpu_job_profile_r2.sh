#!/bin/bash
# round-2 evidence: launch list of one C5 step, ncu --set full of the top kernels, sanitizer logs
set -u
mkdir -p gpurun_out
BARGS="--tiles 8 --steps 2 --warmup 1 --configs none --no-cpu-baseline --no-e2e --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(k_|.*cub).*' -c 900 --csv --log-file gpurun_out/launches.csv python bench.py $BARGS > gpurun_out/launches.log 2>&1; echo "launch list rc=$?"
for k in k_patch_fit_warp k_apply_assign k_a1_search; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 10 -c 1 -f -o gpurun_out/prof_$k python bench.py --tiles 4 --steps 1 --warmup 1 --configs none --no-cpu-baseline --no-e2e --no-graph --streams 1 > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_desc_nn_tc -s 2 -c 1 -f -o gpurun_out/prof_k_desc_nn_tc python tools/bench_desc.py --n 151552 --m 262144 --d 64 --reps 1 > gpurun_out/ncu_desc.log 2>&1; echo "ncu desc64 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_desc_nn_tc -s 2 -c 1 -f -o gpurun_out/prof_k_desc_nn_tc_d32 python tools/bench_desc.py --n 151552 --m 262144 --d 32 --reps 1 > gpurun_out/ncu_desc32.log 2>&1; echo "ncu desc32 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seg_attention_pool_mma -s 4 -c 1 -f -o gpurun_out/prof_k_seg_attention_pool python tools/run_c3_tile.py 300000 1 > gpurun_out/ncu_pool.log 2>&1; echo "ncu pool rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_desc_gpu.py tests/test_exchange_gpu.py tests/test_pinned_gpu.py tests/test_fine_matching_gpu.py -m gpu -q -x > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_exchange_gpu.py tests/test_pinned_gpu.py -m gpu -q -x -k "copy_kernel or filtering or attention" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/sanitizer_racecheck.log
ls -la gpurun_out/*.ncu-rep
