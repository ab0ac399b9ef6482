#!/bin/bash
# fit-kernel loop: parity tests that touch the fit / ICP kernels, phase timers (debug build), stream-overlap ms per tile
set -u
timeout 600 python -m pytest tests/test_fine_matching_gpu.py tests/test_icp_gpu.py tests/test_rigid_gpu.py tests/test_pipeline_gpu.py tests/test_host_api_gpu.py -x -q 2>&1 | tail -4
[ -f fusion4landslide_b200/libf4l_b200_dbg.so ] && F4L_LIB=$PWD/fusion4landslide_b200/libf4l_b200_dbg.so python scratch/stats.py 2>&1 | tail -12
NT=8 python tools/exp_fit_overlap.py 2>&1 | tail -4
