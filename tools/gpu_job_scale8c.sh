#!/bin/bash
set -u
mkdir -p gpurun_out
N=8
for extra in "--no-gather" "--exchange nccl"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $extra > gpurun_out/bench_n8x.json 2> gpurun_out/bench_n8x.err; echo "N=8 [$extra] rc=$?"
python -c "
import json,sys; s=open('gpurun_out/bench_n8x.json').read(); d=json.loads(s[s.index('{\"metric'):]); print('value %.1fM pts/s  ms %.3f' % (d['value']/1e6, d['ms_per_step'])); print(d.get('breakdown'))"
done
