#!/usr/bin/env python
"""Driver for ncu captures of the A1 / K-a kernels: median resolution (+ optional kNN with indices) of one
synthetic TLS tile pair.   python tools/prof_a1.py --n 4000000 [--knn]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4_000_000)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--knn", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
d = synth.make_tile(a.n, seed=1, device=dev, patch_pts=256)
for _ in range(a.reps):
    m = ops.median_resolution(d["src"], d["tgt"])
    if a.knn:
        ops.knn_grid(d["src"], d["src"], 2)
        ops.knn_grid(d["src"], d["tgt"], 1)
torch.cuda.synchronize()
print("median resolution", float(m))
