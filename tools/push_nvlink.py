#!/usr/bin/env python
"""One process, two GPUs: the copy kernel of the displacement-field exchange (f4l_peer_push) from cuda:0 into a buffer on
cuda:1 over NVLink -- duration with CUDA events (GB/s per peer) and, under
    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum -k regex:k_peer_push ...
the NVLink byte counters of one launch.   python tools/push_nvlink.py [rows] [peers]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import _lib, ops  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 781_250
n_peers = int(sys.argv[2]) if len(sys.argv) > 2 else 1
assert torch.cuda.device_count() >= 2, "needs 2 GPUs"
dev0 = torch.device("cuda:0")
with torch.cuda.device(0):
    for d in range(1, min(torch.cuda.device_count(), n_peers + 1)):
        _lib.check(_lib.lib().f4l_peer_enable_access(d), "f4l_peer_enable_access")
    src = torch.randn((rows, 6), device=dev0)
    count = torch.tensor([rows, 0, 0, 0], dtype=torch.int32, device=dev0)
    remotes = [torch.zeros((rows, 6), device="cuda:%d" % (1 + (i % (torch.cuda.device_count() - 1)))) for i in range(n_peers)]
    ptrs = [r.data_ptr() for r in remotes]
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)
    ops.peer_push(src, count, ptrs)
    torch.cuda.synchronize(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ops.peer_push(src, count, ptrs)
    e1.record()
    torch.cuda.synchronize(0)
    ms = e0.elapsed_time(e1) / reps
    ok = all(torch.equal(r.to(dev0), src) for r in remotes)
    nbytes = rows * 24 * n_peers
    print("k_peer_push: %d rows x 24 B to %d peer buffer(s): %.3f ms per launch, %.1f GB/s NVLink egress, copies equal: %s"
          % (rows, n_peers, ms, nbytes / ms / 1e6, ok))
