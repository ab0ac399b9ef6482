#!/usr/bin/env python
"""Stand-alone kernel micro-benchmarks against the HBM roofline (MEASURED_PEAKS.json): K-a grid kNN,
K-d segmented Kabsch reduction, K-f transform apply, K-c rigidity, at tile scale.  Inputs exceed L2
(or L2 is flushed between iterations); CUDA events; algorithmic bytes as in DESIGN.md section 4.
    python tools/bench_kernels.py [--n 4000000] [--patch 256]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import _lib, ops, synth  # noqa: E402

LAST_KERNELS = {}


def timed(fn, reps, flush):
    """Median over reps of the SUM of the library's in-stream per-kernel times (CUDA events recorded on the
    launching stream around every kernel of the call), L2 flushed before each rep.  The per-kernel split of the
    last rep is left in LAST_KERNELS; host launch overhead is excluded (it dominates at these sizes)."""
    L = _lib.lib()
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()                       # 256 MB write: evicts the 126 MB L2
        torch.cuda.synchronize()
        L.f4l_profile_reset()
        L.f4l_profile_enable(1)
        fn()
        torch.cuda.synchronize()
        L.f4l_profile_enable(0)
        tab = _lib.profile_table()
        ts.append(sum(v[0] for v in tab.values()))
        LAST_KERNELS.clear()
        LAST_KERNELS.update({k: round(v[0], 5) for k, v in tab.items()})
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4_000_000)
    ap.add_argument("--patch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--only", default="", help="'rigid': K-d / K-f / K-c only (skip the kNN / voxel part)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    K, Q = a.n, a.n // a.patch
    src = (torch.rand((K, 3), device=dev, generator=g) * 100).contiguous()
    tgt = (src + 0.05 * torch.randn((K, 3), device=dev, generator=g)).contiguous()
    w = torch.rand((K,), device=dev, generator=g)
    ptr = torch.arange(0, K + 1, a.patch, dtype=torch.int32, device=dev)
    Q = ptr.numel() - 1
    out = {}

    def rec(name, ms, nbytes, extra=None):
        gbs = nbytes / ms / 1e6
        out[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak,
                     "kernels_ms": dict(LAST_KERNELS)}
        if extra:
            out[name].update(extra)

    ms = timed(lambda: ops.segmented_kabsch(src, tgt, ptr, eps=1e-6, variant=0), a.reps, flush)
    rec("k_segmented_kabsch (packed, no weights)", ms, 24 * K + 4 * (Q + 1) + 48 * Q)
    ms = timed(lambda: ops.segmented_kabsch(src, tgt, ptr, w=w, eps=1e-7, variant=1, want_res=True), a.reps, flush)
    rec("k_segmented_kabsch (weights + residuals: 2 passes)", ms, 2 * 24 * K + 4 * K + 4 * K + 48 * Q)
    R, t, _ = ops.segmented_kabsch(src, tgt, ptr, eps=1e-6, variant=0)
    T = torch.zeros((Q, 4, 4), device=dev)
    T[:, :3, :3] = R
    T[:, :3, 3] = t
    T[:, 3, 3] = 1
    ms = timed(lambda: ops.apply_transforms(src, ptr, T), a.reps, flush)
    rec("k_apply_transforms (dvf + mag)", ms, 40 * K + 64 * Q)
    ms = timed(lambda: ops.rigidity_check(src, tgt, ptr, 0.5), a.reps, flush)
    rec("k_rigidity (K^2/2 pairs per patch)", ms, 24 * K + 8 * Q, {"pair_evals_per_s": Q * a.patch * (a.patch - 1) / 2 / ms * 1e3})
    # K-a on a synthetic TLS tile pair
    for n in (() if a.only == "rigid" else (1_000_000, a.n)):
        d = synth.make_tile(n, seed=1, device=dev, patch_pts=a.patch)
        s, tg = d["src"], d["tgt"]
        ms = timed(lambda: ops.knn_grid(s, s, 2), a.reps, flush)
        rec("f4l_knn_grid self k=2 N=%d (bin + search)" % n, ms, 12 * n + 12 * n + 16 * n, {"queries_per_s": n / ms * 1e3})
        ms = timed(lambda: ops.knn_grid(s, tg, 1), a.reps, flush)
        rec("f4l_knn_grid cross k=1 N=M=%d (bin + search)" % n, ms, 12 * n + 12 * n + 8 * n, {"queries_per_s": n / ms * 1e3})
        ms = timed(lambda: ops.median_resolution(s, tg), a.reps, flush)
        rec("f4l_median_resolution N=M=%d" % n, ms, 2 * (24 * n + 16 * n), {"points_per_s": 2 * n / ms * 1e3})
        s64 = s.double().contiguous()
        nv = ops.voxel_downsample(s64, 0.1).shape[0]
        ms = timed(lambda: ops.voxel_downsample(s64, 0.1, want_map=True), a.reps, flush)
        # 24 B per point in, 4 B map out, 24 B per voxel out (the sort passes are extra traffic on top)
        rec("f4l_voxel_downsample N=%d (%d voxels of 0.1 m)" % (n, nv), ms, 28 * n + 24 * nv, {"points_per_s": n / ms * 1e3})
        del d, s64
    print(json.dumps({"hbm_peak_gbs": peak, "kernels": out}, indent=1))


if __name__ == "__main__":
    main()
