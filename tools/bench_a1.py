#!/usr/bin/env python
"""A1 (median resolution) micro-benchmark: per-kernel in-stream times at tile scale and beyond, L2 flushed.
    python tools/bench_a1.py [--n 781250 4000000 16000000]    (F4L_KNN_CELL_FACTOR=x to sweep the cell size)"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import _lib, ops, synth  # noqa: E402
from tools.bench_kernels import LAST_KERNELS, timed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[781_250, 4_000_000, 16_000_000])
ap.add_argument("--reps", type=int, default=7)
a = ap.parse_args()
dev = torch.device("cuda:0")
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {"hbm_peak_gbs": peak, "cell_factor": os.environ.get("F4L_KNN_CELL_FACTOR", "default"), "sizes": {}}
for n in a.n:
    d = synth.make_tile(n, seed=1, device=dev, patch_pts=256)
    s, tg = d["src"], d["tgt"]
    ms = timed(lambda: ops.median_resolution(s, tg), a.reps, flush)
    k = dict(LAST_KERNELS)
    nn = s.shape[0] + tg.shape[0]
    rec = {"ms_total": ms, "points_per_s": nn / ms * 1e3, "kernels_ms": k, "median_resolution": float(ops.median_resolution(s, tg))}
    sk = "k_a1_search_tiled" if "k_a1_search_tiled" in k else "k_a1_search"
    if sk in k:
        rec["search_frac_of_hbm (40 B/pt)"] = 40 * nn / (k[sk] * 1e-3) / 1e9 / peak
        rec["scatter_frac_of_hbm (28 B/pt)"] = 28 * nn / (k["k_a1_scatter"] * 1e-3) / 1e9 / peak
        rec["pipeline_frac_of_hbm (80 B/pt)"] = 80 * nn / (ms * 1e-3) / 1e9 / peak
    out["sizes"][str(n)] = rec
    del d, s, tg
print(json.dumps(out, indent=1))
