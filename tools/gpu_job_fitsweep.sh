#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "1 0" "2 0" "4 0" "8 0" "16 0" "4 3" "8 3" "16 3" "8 2" "16 2"; do
  set -- $cfg
  python bench.py --configs none --no-cpu-baseline --no-e2e --fit-groups $1 --fit-ctas $2 --steps 5 > gpurun_out/fs_$1_$2.json 2>gpurun_out/fs.err || tail -3 gpurun_out/fs.err
  python -c "
import json; s=open('gpurun_out/fs_$1_$2.json').read(); d=json.loads(s[s.index('{\"'):]); print('groups $1 ctas $2: %.1f M pts/s  %.2f ms' % (d['value']/1e6, d['ms_per_step']))"
done
