#!/usr/bin/env python
"""Experiment: what bounds the host-buffer (e2e) path?  16 C5 tiles: HostPipeline with / without the sparse rows,
and the raw PCIe copies alone (H2D inputs, D2H outputs, both at once)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import pipeline, synth

dev = torch.device("cuda:0")
NT = int(os.environ.get("NT", 16))
tiles = []
for s in range(NT):
    d = synth.make_tile(781_250, seed=s, patch_pts=256, device=dev)
    tiles.append(pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"]))
host_tiles = [pipeline.HostTile(t) for t in tiles]


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for ws in (True, False):
    kw = {}
    if os.environ.get("SPARSE_ONCE") and ws:
        kw["sparse_once"] = True
    hp = pipeline.HostPipeline(host_tiles, None, dev, n_streams=4, want_sparse=ws, **kw)
    res, bi, bo = hp.run()
    ms = timeit(hp.run)
    print("HostPipeline want_sparse=%s %s: %.2f ms per %d tiles (%.1f M pts/s)  h2d %.2f GB d2h %.2f GB" % (ws, kw, ms, NT, NT * 781250 / ms / 1e3, bi / 1e9, bo / 1e9), flush=True)
    del hp

streams = [torch.cuda.Stream() for _ in range(4)]
dbuf = [{k: v.to(dev) for k, v in ht.t.items()} for ht in host_tiles]
hout = [torch.empty((781250 * 3, 6), dtype=torch.float32).pin_memory() for _ in range(NT)]
dout = [torch.empty((781250 * 3, 6), dtype=torch.float32, device=dev) for _ in range(4)]
nb_in = sum(ht.nbytes() for ht in host_tiles)


def h2d():
    for i, ht in enumerate(host_tiles):
        with torch.cuda.stream(streams[i % 4]):
            for k, v in ht.t.items():
                dbuf[i][k].copy_(v, non_blocking=True)


def d2h():
    for i in range(NT):
        with torch.cuda.stream(streams[i % 4]):
            hout[i].copy_(dout[i % 4], non_blocking=True)


def both():
    h2d(); d2h()


for name, fn, nb in (("H2D", h2d, nb_in), ("D2H", d2h, NT * 781250 * 3 * 24), ("both", both, nb_in + NT * 781250 * 3 * 24)):
    ms = timeit(fn)
    print("%s alone: %.2f ms, %.1f GB/s" % (name, ms, nb / ms / 1e6), flush=True)
