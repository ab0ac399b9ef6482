#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
run() { tag=$1; shift
  timeout 600 env "$@" python bench.py --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "$tag rc=$?"
  tail -2 gpurun_out/bench_$tag.err | cut -c1-300
  python -c "
import json; s=open('gpurun_out/bench_$tag.json').read(); d=json.loads(s[s.index('{'):]); print('$tag: value %.1fM pts/s  ms %.2f launches %d' % (d['value']/1e6, d['ms_per_step'], d['gpu_launches']))
for k,v in list(d['kernels'].items())[:4]: print('   %-22s %.4f ms x%d share %.3f' % (k, v['ms_avg'], v['launches'], v['share']))"
}
run base A=1
run phased F4L_TILE_ORDER=phased
BENCH_ARGS="--streams 8" run phased8 F4L_TILE_ORDER=phased
