#!/bin/bash
# fastest loop: fine-matching / pipeline parity tests + the bench line without CPU baseline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest ${TESTS:-tests} -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 600 python bench.py --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err | cut -c1-300
python -c "
import json; s=open('gpurun_out/bench.json').read(); d=json.loads(s[s.index('{'):]); print('value %.1fM pts/s  ms %.2f  e2e %s launches %d' % (d['value']/1e6, d['ms_per_step'], d['e2e'] and '%.1fM (%.1f ms)' % (d['e2e']['value']/1e6, d['e2e']['ms_per_step']), d['gpu_launches']))
for k,v in list(d['kernels'].items())[:12]: print('   %-22s %.4f ms x%d share %.3f' % (k, v['ms_avg'], v['launches'], v['share']))"
