#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python tools/bench_kernels.py > gpurun_out/bench_kernels.json 2> gpurun_out/bench_kernels.err; echo "kernels rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_kernels.json'))
for k,v in d['kernels'].items(): print('%-60s %8.3f ms %8.1f GB/s %5.1f%%' % (k, v['ms'], v['GB/s'], 100*v['frac_of_measured_hbm']))"
timeout 600 python tools/exp_fit_overlap.py 2>&1 | tail -6
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; s=open('gpurun_out/bench.json').read(); d=json.loads(s[s.index('{'):]); print('value %.1fM pts/s  ms %.2f launches %d' % (d['value']/1e6, d['ms_per_step'], d['gpu_launches']))
for k,v in list(d['kernels'].items())[:12]: print('   %-22s %.4f ms x%d share %.3f' % (k, v['ms_avg'], v['launches'], v['share']))"
