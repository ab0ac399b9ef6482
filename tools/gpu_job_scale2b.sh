#!/bin/bash
# 2-GPU check of the pipelined push exchange (collective on the exchange stream) + exchange tests
set -u
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m pytest tests/test_exchange_gpu.py -x -q -m gpu 2>&1 | tail -2
for extra in "--exchange push" "--exchange push --no-gather"; do
  tag=$(echo $extra | tr -d ' -')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --configs none $extra > gpurun_out/bench_n${N}_${tag}.json 2> gpurun_out/bench_n${N}_${tag}.err; echo "N=$N [$extra] rc=$?"
  tail -3 gpurun_out/bench_n${N}_${tag}.err | cut -c1-300
  python -c "
import json,sys; s=open('gpurun_out/bench_n${N}_${tag}.json').read(); d=json.loads(s[s.index('{\"metric'):]); print('value %.1fM pts/s  ms %.3f' % (d['value']/1e6, d['ms_per_step'])); print(d.get('breakdown'))"
done
