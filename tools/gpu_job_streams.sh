#!/bin/bash
set -u
mkdir -p gpurun_out
for s in 1 2 4 8; do
  timeout 600 python bench.py --streams $s --no-cpu-baseline > gpurun_out/bench_streams_$s.json 2> gpurun_out/bench_streams_$s.err; echo "streams=$s rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_streams_$s.json')); print('value %.1fM pts/s  ms %.2f  e2e %.1fM (%.1f ms)' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']))"
  tail -3 gpurun_out/bench_streams_$s.err
done
