#!/bin/bash
# DIPs front-end loop: parity tests, then `bench.py --workload dips` (its CPU leg is the oracle port)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dips_gpu.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload dips ${DIPS_ARGS:-} > gpurun_out/bench_dips.json 2> gpurun_out/bench_dips.err; tail -3 gpurun_out/bench_dips.err
python -c "
import json; d=json.load(open('gpurun_out/bench_dips.json'))
print('ms %.3f  patches/s %.1fM  hbm frac %.3f  cpu %.0f/s  nb mean %.0f' % (d['ms_per_step'], d['value']/1e6, d['roofline']['frac'], d.get('cpu_baseline', {}).get('value', 0), d['config']['neighbours_mean']))
print({k: v['ms_total'] for k, v in d['kernels'].items()})"
if [ -n "${NCU:-}" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dips_patches -s 1 -c 1 -f -o gpurun_out/prof_k_dips_patches python bench.py --workload dips --dips-pts 200000 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_dips.log 2>&1; echo "ncu rc=$?"
fi
