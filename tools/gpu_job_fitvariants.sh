#!/bin/bash
# A/B of fit-kernel occupancy variants (FITW_WARPS, FITW_MIN_BLOCKS, WICP_CAP) built as separate libraries under scratch/variants/
set -u
mkdir -p gpurun_out
for lib in scratch/variants/libf4l_*.so; do
  name=$(basename $lib .so)
  F4L_LIB=$PWD/$lib python bench.py --configs none --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/fv_$name.json 2>gpurun_out/fv.err || tail -3 gpurun_out/fv.err
  python -c "
import json; s=open('gpurun_out/fv_$name.json').read(); d=json.loads(s[s.index('{\"'):]); k=[x for x in d.get('kernels',[]) if 'fit' in str(x)]; print('$name: %.1f M pts/s  %.2f ms' % (d['value']/1e6, d['ms_per_step']), k[:3])"
done
