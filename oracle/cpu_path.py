"""TEST INFRASTRUCTURE / CPU ARM -- the hot path of one tile on host cores, as the reference runs it.

Used by bench.py's `cpu_baseline` leg and `--impl reference` arm, and by __graft_entry__.smoke()
as the checker.  It is the oracle port (kind "port"): Open3D / hnswlib / faiss are not installable
here, so the reference's own per-patch loop (base.py:3254-3438) is executed through the
restatements in oracle/fine_matching.py with scipy's cKDTree in the role of Open3D's KD-tree.
The reference runs this loop serially in one Python thread; here the patch pairs are spread over
`workers` forked processes so the arm uses every host core it can.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import fine_matching as ofm
from . import knn as oknn

_G = {}


def _work(args):
    lo, hi = args
    g = _G
    o = ofm.fine_matching(g["src"], g["tgt"], g["corr3d"], None, g["spt_src"][lo:hi], g["spt_tgt"][lo:hi], g["prm"])
    rows = sum(0 if d is None else d.shape[0] for d in o["dense"])
    return lo, hi, rows, o["status"], o["T"], o["iters"], o["K"]


def run_tile(src, tgt, corr3d, spt_src, spt_tgt, workers=None, max_pairs=None, **fine_kwargs):
    """Median resolution (A1) + fine matching of one tile on the CPU.

    max_pairs bounds the sample (first max_pairs patch pairs).  Returns dict(seconds, seconds_median,
    seconds_fine, src_points (points of the processed source patches), dense_rows, status, T, workers)."""
    workers = workers or os.cpu_count() or 1
    t0 = time.perf_counter()
    med = oknn.median_resolution(src, tgt)                      # kd-tree k=2 self query x2, all cores
    t1 = time.perf_counter()
    Q = len(spt_src) if max_pairs is None else min(max_pairs, len(spt_src))
    prm = ofm.FineParams(median_max_resolution=med, **fine_kwargs)
    _G.update(src=src, tgt=tgt, corr3d=corr3d, spt_src=spt_src, spt_tgt=spt_tgt, prm=prm)
    chunk = max(1, Q // (workers * 4))
    jobs = [(lo, min(lo + chunk, Q)) for lo in range(0, Q, chunk)]
    status = np.zeros(Q, np.int8)
    T = np.tile(np.eye(4, dtype=np.float32), (Q, 1, 1))
    iters = np.zeros(Q, np.int64)
    K = np.zeros(Q, np.int64)
    rows = 0
    if workers > 1 and len(jobs) > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            it = pool.imap_unordered(_work, jobs)
            results = list(it)
    else:
        results = [_work(j) for j in jobs]
    for lo, hi, r, st, Tq, itq, Kq in results:
        rows += r
        status[lo:hi] = st
        T[lo:hi] = Tq
        iters[lo:hi] = itq
        K[lo:hi] = Kq
    t2 = time.perf_counter()
    pts = int(sum(len(spt_src[q]) for q in range(Q)))
    return dict(seconds=t2 - t0, seconds_median=t1 - t0, seconds_fine=t2 - t1, src_points=pts,
                dense_rows=rows, status=status, T=T, iters=iters, K=K, workers=workers, median_resolution=med, pairs=Q)
