#!/bin/bash
# K-b ablations: full kernel / TMEM loads only (dbg 1) / MMA only (dbg 2), D = 32 and 64, with the SM clock sampled
set -u
mkdir -p gpurun_out
for dbg in 0 1 2; do
 for d in 64 32; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 50 > gpurun_out/clk_${d}_$dbg.txt &
  SMI=$!
  F4L_DESC_DBG=$dbg timeout 300 python tools/bench_desc.py --d $d --n 524288 --m 524288 --reps 6 > gpurun_out/bench_desc_${d}_dbg$dbg.json 2> gpurun_out/bench_desc_${d}_dbg$dbg.err
  kill $SMI
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_desc_${d}_dbg$dbg.json'))
rows=[l.split(',') for l in open('gpurun_out/clk_${d}_$dbg.txt') if ',' in l]
clk=sorted(float(r[0]) for r in rows); pw=max(float(r[1]) for r in rows)
print('dbg=$dbg D=$d  tc %.1f ms  %.0f TFLOP/s useful   clk min/med/max %s  power max %.0f W' % (d['kernels']['k_desc_nn_tc']['ms_avg'], d.get('tc_kernel_tflops',0), (clk[0], clk[len(clk)//2], clk[-1]), pw))
PY
 done
done
