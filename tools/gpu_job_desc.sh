#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_desc_gpu.py tests/test_piecewise_gpu.py -m gpu -q > gpurun_out/pytest_desc.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_desc.log | cut -c1-300
for dbg in 0 2; do
 for d in 64 32; do
  F4L_DESC_DBG=$dbg timeout 300 python tools/bench_desc.py --d $d --n 524288 --m 524288 > gpurun_out/bench_desc_${d}_dbg$dbg.json 2> gpurun_out/bench_desc_${d}_dbg$dbg.err; echo "dbg=$dbg d=$d rc=$?"; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_desc_${d}_dbg$dbg.json')); print(d['kernels'], d.get('tc_kernel_tflops'), d['matched_to_first_half'])"
 done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_desc_nn_tc -s 1 -c 1 -f -o gpurun_out/prof_k_desc_nn_tc python tools/bench_desc.py --n 151552 --m 131072 --reps 1 > gpurun_out/ncu_desc.log 2>&1; echo "ncu rc=$?"
