#!/bin/bash
set -u
mkdir -p gpurun_out
for mode in "" "--no-graph"; do
  timeout 600 python bench.py --no-cpu-baseline --no-e2e $mode > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "mode=[$mode] rc=$?"
  tail -2 gpurun_out/bench_g.err
  python -c "
import json; s=open('gpurun_out/bench_g.json').read(); d=json.loads(s[s.index('{'):]); print('value %.1fM pts/s  ms %.2f launches %d graph %s' % (d['value']/1e6, d['ms_per_step'], d['gpu_launches'], d['config']['cuda_graph']))"
done
