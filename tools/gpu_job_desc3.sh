#!/bin/bash
set -u
for dbg in ${DBGS:-4 5}; do for d in 64 32; do
  F4L_DESC_DBG=$dbg timeout 300 python tools/bench_desc.py --d $d --n 524288 --m 524288 --reps 2 2>&1 | grep "k_desc_nn_ts" | head -2
done; done
