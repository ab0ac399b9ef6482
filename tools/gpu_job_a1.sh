#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_fine_matching_gpu.py -x -q 2>&1 | tail -5
for f in ${FACTORS:-1.5}; do
  F4L_KNN_CELL_FACTOR=$f timeout 300 python tools/bench_a1.py > gpurun_out/a1_f$f.json 2> gpurun_out/a1_f$f.err || tail -5 gpurun_out/a1_f$f.err
  python - <<EOF
import json
d=json.load(open('gpurun_out/a1_f$f.json'))
for n,r in d['sizes'].items():
    print('factor $f n=%s total %.3f ms  %.2f Gpts/s  search %.3f (%.3f of hbm) ' % (n, r['ms_total'], r['points_per_s']/1e9, r['kernels_ms'].get('k_a1_search',0), r.get('search_frac_of_hbm (40 B/pt)',0)), r['kernels_ms'])
EOF
done
