#!/bin/bash
# ncu --set full capture of the kNN search kernel (second epoch of the second repetition) + launch list
set -u
mkdir -p gpurun_out
N=${N:-4000000}
for k in ${KERNELS:-k_grid_search}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-1} -c 1 -f -o gpurun_out/prof_$k python tools/prof_a1.py --n $N ${EXTRA:-} > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
