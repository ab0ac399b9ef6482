#!/bin/bash
# ICP fragility flag: tests + a short bench (C5 only) to check the parity record and the step time
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_icp_gpu.py tests/test_fine_matching_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12
python bench.py --configs none --no-e2e --steps 10 > gpurun_out/bench_flag.json 2>gpurun_out/bench_flag.err || tail -3 gpurun_out/bench_flag.err
python -c "
import json; s=open('gpurun_out/bench_flag.json').read(); d=json.loads(s[s.index('{\"'):]); print('%.1f M pts/s  %.2f ms' % (d['value']/1e6, d['ms_per_step'])); print(d.get('parity'))"
