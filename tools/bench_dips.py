#!/usr/bin/env python
"""DIPs patch front-end (f4l_dips_build + f4l_dips_patches) at tile scale next to the CPU restatement.

One C3-shaped tile (625 k points per epoch at 0.1 m spacing, feature radius sqrt(3)*10*resolution = 1.73 m, about
940 neighbours per point), every point a query: patches/s, the output-write fraction of the HBM roofline
(algorithmic bytes = 24 B in + 3*256*4 B out per point), the per-kernel split, and oracle/dips.py (numpy + cKDTree,
one core, the reference's own per-point loop shape) on a bounded sample.
    python tools/bench_dips.py [--n 625000] [--cpu-queries 300]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import _lib, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=625_000)
    ap.add_argument("--batch", type=int, default=125_000, help="queries per launch (output batch = 3 KB per query)")
    ap.add_argument("--cpu-queries", type=int, default=300)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    d = synth.make_tile(a.n, seed=3, device=dev)
    ref = d["src"].double().contiguous()
    radius = float(np.sqrt(3) * 10 * 0.1)
    L = _lib.lib()
    out = torch.empty((a.batch, 3, 256), dtype=torch.float32, device=dev)

    def run():
        index = ops.DipsIndex(ref, radius)
        cnt = []
        for off in range(0, a.n, a.batch):
            q = ref[off:off + a.batch]
            p, c = ops.dips_patches(index, q, 256, seed=off, out=out[:q.shape[0]])
            cnt.append(c)
        return torch.cat(cnt)

    cnt = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    L.f4l_profile_reset(); L.f4l_profile_enable(1)
    run()
    torch.cuda.synchronize()
    L.f4l_profile_enable(0)
    kern = {k: {"ms_total": round(v[0], 4), "launches": v[1]} for k, v in _lib.profile_table().items()}
    alg_bytes = a.n * (24 + 3 * 256 * 4)
    res = {
        "workload": "DIPs front-end, 1 tile of %d points (both the cloud and the queries), radius %.3f m" % (a.n, radius),
        "neighbours_mean": float(cnt.float().mean()), "neighbours_max": int(cnt.max()),
        "ms_per_tile_epoch": ms, "patches_per_s": a.n / (ms * 1e-3),
        "algorithmic_bytes": alg_bytes, "GB/s": alg_bytes / (ms * 1e-3) / 1e9,
        "frac_of_measured_hbm": alg_bytes / (ms * 1e-3) / 1e9 / peak, "hbm_peak_gbs": peak,
        "kernels": kern,
    }
    # CPU: oracle/dips.py on a bounded sample (test infrastructure used as the reported baseline only)
    from oracle import dips as odips
    refn = ref.cpu().numpy()
    rng = np.random.default_rng(0)
    pick = rng.choice(a.n, a.cpu_queries, replace=False)
    inds = np.stack([rng.permutation(2048)[:256] % 256 for _ in pick])
    t0 = time.perf_counter()
    from scipy.spatial import cKDTree
    tree = cKDTree(refn)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i, q in enumerate(refn[pick]):
        pa, _, _ = odips.extract_all(q, tree, refn, radius)
        odips.sample(pa, inds[i] % max(pa.shape[0], 256))
    t_q = time.perf_counter() - t0
    res["cpu_baseline"] = {"kind": "port", "cores": 1, "sample": "%d queries of the same tile (oracle/dips.py, numpy + cKDTree; tree build %.2f s excluded)" % (a.cpu_queries, t_build),
                           "patches_per_s": a.cpu_queries / t_q}
    res["speedup_vs_cpu_port"] = res["patches_per_s"] / res["cpu_baseline"]["patches_per_s"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
