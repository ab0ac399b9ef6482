#!/bin/bash
# K-d: tests, racecheck, micro-benchmark at three sizes, one ncu capture of k_kabsch_fused
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rigid_gpu.py tests/test_host_api_gpu.py tests/test_rgb_guided_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_rigid_gpu.py -m gpu -x -q -k "kabsch" > gpurun_out/sanitizer_racecheck_rigid.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_rigid.txt
for n in 1000000 4000000 16000000; do
  python tools/bench_kernels.py --only rigid --n $n > gpurun_out/kab_fused_$n.json 2>gpurun_out/kab.err || tail -3 gpurun_out/kab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/kab_fused_$n.json"))["kernels"]
for k,v in d.items():
    if "kabsch" in k or "apply" in k: print("$n", k, "%.4f ms  %.0f GB/s  frac %.3f" % (v["ms"], v["GB/s"], v["frac_of_measured_hbm"]), v["kernels_ms"])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_kabsch_fused -s 1 -c 1 -f -o gpurun_out/prof_k_kabsch_fused python tools/bench_kernels.py --only rigid --reps 1 --n 16000000 > gpurun_out/ncu_kab.log 2>&1; echo "ncu rc=$?"
