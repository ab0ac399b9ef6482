#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke, the reference arm, the bench line
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python -c "
import json; s=open('gpurun_out/bench_ref.json').read(); d=json.loads(s[s.index('{\"'):]); print('reference arm: %.3f M pts/s, %.1f ms/step, cores %s' % (d['value']/1e6, d['ms_per_step'], d['cpu_baseline']['cores']))"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench.json
