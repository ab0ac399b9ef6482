#!/bin/bash
set -u
mkdir -p gpurun_out
[ -n "${SKIP_TESTS:-}" ] || timeout 600 python -m pytest tests/test_desc_gpu.py -m gpu -q -x > gpurun_out/pytest_desc.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_desc.log | cut -c1-300
for dbg in ${DBGS:-0 2}; do
 for d in 64 32; do
  F4L_DESC_DBG=$dbg timeout 300 python tools/bench_desc.py --d $d --n 524288 --m 524288 --reps 6 > gpurun_out/bench_desc_${d}_dbg$dbg.json 2> gpurun_out/bench_desc_${d}_dbg$dbg.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_desc_${d}_dbg$dbg.json'))
print('dbg=$dbg D=$d  tc %.1f ms  %.0f TFLOP/s useful  match %.3f' % (d['kernels']['k_desc_nn_tc']['ms_avg'], d.get('tc_kernel_tflops',0), d['matched_to_first_half']))
PY
 done
done
