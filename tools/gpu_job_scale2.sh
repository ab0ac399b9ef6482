#!/bin/bash
# N-GPU check: exchange tests, then bench at N with the fused exchange and with the NCCL all-gather baseline
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
N=${NGPU:-2}
timeout 600 python -m pytest tests/test_exchange_gpu.py -x -q -m gpu > gpurun_out/pytest_exchange.log 2>&1; echo "pytest exchange rc=$?"
tail -15 gpurun_out/pytest_exchange.log | cut -c1-300
for extra in "--exchange push" "--exchange fused" ${EXTRA_MODES:-}; do
  tag=$(echo $extra | tr -d ' -')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $extra > gpurun_out/bench_n${N}_${tag}.json 2> gpurun_out/bench_n${N}_${tag}.err; echo "N=$N [$extra] rc=$?"
  tail -3 gpurun_out/bench_n${N}_${tag}.err | cut -c1-300
  python -c "
import json,sys; s=open('gpurun_out/bench_n${N}_${tag}.json').read(); d=json.loads(s[s.index('{\"metric'):]); print('value %.1fM pts/s  ms %.2f' % (d['value']/1e6, d['ms_per_step'])); print(d.get('breakdown'))"
done
