#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "per-tile 1 0 4" "batched 1 0 4" "batched 2 0 4" "per-tile 1 0 3" "per-tile 1 0 8" "batched 1 0 8"; do
  set -- $cfg
  python bench.py --tiles 8 --configs none --no-cpu-baseline --no-e2e --fit $1 --fit-groups $2 --fit-ctas $3 --streams $4 --steps 10 > gpurun_out/f8.json 2>gpurun_out/f8.err || tail -3 gpurun_out/f8.err
  python -c "
import json; s=open('gpurun_out/f8.json').read(); d=json.loads(s[s.index('{\"'):]); print('8 tiles, fit $1 groups $2 streams $4: %.1f M pts/s  %.3f ms' % (d['value']/1e6, d['ms_per_step']))"
done
