#!/bin/bash
# A/B of library variants on the bench (device-resident value only).  VARIANTS="base c184" ; EXTRA="--streams 4"
set -u
mkdir -p gpurun_out
for v in ${VARIANTS:-base}; do
  lib=""; [ "$v" != base ] && lib=$PWD/fusion4landslide_b200/libf4l_b200_$v.so
  if [ -n "${TESTS:-}" ]; then F4L_LIB=$lib timeout 600 python -m pytest $TESTS -x -q 2>&1 | tail -3; fi
  for ex in "${EXTRA:-}"; do
    F4L_LIB=$lib timeout 600 python bench.py --no-cpu-baseline --no-e2e $ex > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo "$v rc=$?"
    python -c "
import json,sys; s=open('gpurun_out/ab_$v.json').read(); d=json.loads(s[s.index('{'):]); print('$v $ex: value %.1fM pts/s  ms %.2f' % (d['value']/1e6, d['ms_per_step']))
for k,v in list(d['kernels'].items())[:6]: print('   %-22s %.4f ms x%d share %.3f' % (k, v['ms_avg'], v['launches'], v['share']))"
  done
done
