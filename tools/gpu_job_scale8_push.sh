#!/bin/bash
set -u
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "N=8 rc=$?"
tail -3 gpurun_out/bench_n8.err | cut -c1-300
python -c "
import json,sys; s=open('gpurun_out/bench_n8.json').read(); d=json.loads(s[s.index('{\"metric'):]); print('value %.1fM pts/s  ms %.3f  e2e %s' % (d['value']/1e6, d['ms_per_step'], d.get('e2e') and (d['e2e']['value']/1e6, d['e2e']['ms_per_step']))); print(d.get('breakdown'))"
