#!/bin/bash
set -u
mkdir -p gpurun_out
for t in 1 2 3; do
  python bench.py --configs C3 --no-cpu-baseline --no-e2e --steps 3 --c2f-threads $t > gpurun_out/c3_t$t.json 2>gpurun_out/c3.err || tail -5 gpurun_out/c3.err
  python -c "
import json; s=open('gpurun_out/c3_t$t.json').read(); d=json.loads(s[s.index('{\"'):]); c=d['configs']['C3']; print('threads $t: %.2f M pts/s  %.1f ms/step rows %d' % (c['value']/1e6, c['ms_per_step'], c['dvf_points_per_step']))"
done
