#!/bin/bash
# experiment: same 50 M-point job cut into fewer, larger tiles (upper bound of what cross-tile batching can give)
set -u
mkdir -p gpurun_out
for cfg in "64 781250" "16 3125000" "4 12500000" "1 50000000"; do
  set -- $cfg
  timeout 900 python bench.py --tiles $1 --tile-pts $2 --no-cpu-baseline --no-e2e > gpurun_out/bench_tiles_$1.json 2> gpurun_out/bench_tiles_$1.err; echo "tiles=$1 rc=$?"
  python -c "
import json; s=open('gpurun_out/bench_tiles_$1.json').read(); d=json.loads(s[s.index('{'):]); print('value %.1fM pts/s  ms %.2f launches %d' % (d['value']/1e6, d['ms_per_step'], d['gpu_launches']))
for k,v in list(d['kernels'].items())[:6]: print('   %-22s %.4f ms x%d share %.3f' % (k, v['ms_avg'], v['launches'], v['share']))"
  tail -2 gpurun_out/bench_tiles_$1.err
done
