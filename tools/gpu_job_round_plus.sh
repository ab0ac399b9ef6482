#!/bin/bash
# round job + memcheck / racecheck of the rigid kernels + ncu capture of k_kabsch_fused
set -u
bash tools/gpu_job_round.sh
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_rigid_gpu.py -m gpu -x -q > gpurun_out/sanitizer_memcheck_rigid.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_rigid.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_rigid_gpu.py -m gpu -x -q -k "kabsch" > gpurun_out/sanitizer_racecheck_rigid.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_rigid.txt
python tools/bench_kernels.py --n 4000000 > gpurun_out/bench_kernels_4m.json 2>gpurun_out/bk.err || tail -3 gpurun_out/bk.err
python tools/bench_kernels.py --n 16000000 > gpurun_out/bench_kernels_16m.json 2>gpurun_out/bk.err || tail -3 gpurun_out/bk.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_kabsch_fused -s 1 -c 1 -f -o gpurun_out/prof_k_kabsch_fused python tools/bench_kernels.py --only rigid --reps 1 --n 16000000 > gpurun_out/ncu_kab.log 2>&1; echo "ncu rc=$?"
