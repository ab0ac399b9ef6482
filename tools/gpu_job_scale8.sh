#!/bin/bash
# driver-style scaling check: the default bench command at N GPUs (exactly as the driver launches it), then the reference arm
set -u
mkdir -p gpurun_out
N=${NGPU:-8}
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "N=$N rc=$?"
tail -3 gpurun_out/bench_n${N}.err | cut -c1-300
python -c "
import json,sys; s=open('gpurun_out/bench_n${N}.json').read(); d=json.loads(s[s.index('{\"metric'):]); print('value %.1fM pts/s  ms %.2f e2e %s' % (d['value']/1e6, d['ms_per_step'], d.get('e2e'))); print(d.get('breakdown'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n${N}.json 2> gpurun_out/bench_ref_n${N}.err; echo "ref N=$N rc=$?"
tail -c 600 gpurun_out/bench_ref_n${N}.json
