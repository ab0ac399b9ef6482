#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dips_gpu.py -x -q 2>&1 | tail -3
timeout 600 python tools/bench_dips.py ${DIPS_ARGS:-} > gpurun_out/bench_dips.json 2> gpurun_out/bench_dips.err; tail -3 gpurun_out/bench_dips.err
python -c "
import json; d=json.load(open('gpurun_out/bench_dips.json'))
print('ms %.3f  patches/s %.1fM  hbm frac %.3f  cpu %.0f/s  nb mean %.0f' % (d['ms_per_tile_epoch'], d['patches_per_s']/1e6, d['frac_of_measured_hbm'], d['cpu_baseline']['patches_per_s'], d['neighbours_mean']))
print({k: v['ms_total'] for k, v in d['kernels'].items()})"
if [ -n "${NCU:-}" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dips_patches -s 1 -c 1 -f -o gpurun_out/prof_k_dips_patches python tools/bench_dips.py --n 200000 --batch 100000 --cpu-queries 5 --reps 1 > gpurun_out/ncu_dips.log 2>&1; echo "ncu rc=$?"
fi
