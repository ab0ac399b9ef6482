#!/bin/bash
# A1 tiled-vs-plain search: parity tests, then the A1 micro-benchmark with both kernels
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_fullsize_gpu.py::test_median_resolution_c2_tile tests/test_pipeline_gpu.py -x -q 2>&1 | tail -4
for t in 1 0; do
  F4L_A1_TILED=$t timeout 300 python tools/bench_a1.py > gpurun_out/a1_tiled$t.json 2> gpurun_out/a1_tiled$t.err || tail -5 gpurun_out/a1_tiled$t.err
  python - <<PY
import json
d=json.load(open('gpurun_out/a1_tiled$t.json'))
for n,r in d['sizes'].items():
    print('tiled=$t n=%s total %.3f ms  med %.9f  search frac %.3f ' % (n, r['ms_total'], r['median_resolution'], r.get('search_frac_of_hbm (40 B/pt)',0)), r['kernels_ms'])
PY
done
