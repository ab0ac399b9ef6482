#!/bin/bash
# the driver's bench line (all configs) + the reference arm
set -u
mkdir -p gpurun_out
timeout ${BENCH_TIMEOUT:-900} python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err | cut -c1-300
python tools/show_bench.py gpurun_out/bench.json
