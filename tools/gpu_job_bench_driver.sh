#!/bin/bash
set -u
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?  wall $(( $(date +%s) - T0 )) s"
tail -3 gpurun_out/bench.err | cut -c1-300
python tools/show_bench.py gpurun_out/bench.json
