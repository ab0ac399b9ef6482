#!/bin/bash
# K-d A/B over library variants under scratch/variants/
set -u
mkdir -p gpurun_out
for lib in scratch/variants/*.so; do
F4L_LIB=$PWD/$lib timeout 300 python -m pytest tests/test_rigid_gpu.py -m gpu -x -q 2>&1 | tail -1
for n in 2048 1000000 16000000; do
  F4L_LIB=$PWD/$lib python tools/bench_kernels.py --only rigid --n $n > gpurun_out/kab_ab.json 2>gpurun_out/kab.err || tail -3 gpurun_out/kab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/kab_ab.json"))["kernels"]
for k,v in d.items():
    if "kabsch (packed" in k: print("$lib", "$n", "%.4f ms  frac %.3f" % (v["ms"], v["frac_of_measured_hbm"]), v["kernels_ms"])
PY
done
done
