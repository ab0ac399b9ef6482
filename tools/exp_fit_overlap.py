#!/usr/bin/env python
"""Experiment: do the fine-matching launches of different tiles overlap across streams?  8 tiles of the C5
shape, fine matching only (no A1), on 1/2/4/8 streams; ms per tile."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import ops, pipeline, synth, _lib

dev = torch.device("cuda:0")
NT = int(os.environ.get("NT", 8))
tiles = []
for s in range(NT):
    d = synth.make_tile(781_250, seed=s, patch_pts=256, device=dev)
    tiles.append(pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"]))
cfg = pipeline.FineConfig()
med = torch.tensor([0.05], device=dev)
outs = [None] * NT


def run(streams):
    cur = torch.cuda.current_stream(dev)
    for s in streams:
        s.wait_stream(cur)
    for i, t in enumerate(tiles):
        with torch.cuda.stream(streams[i % len(streams)]):
            outs[i] = ops.fine_matching(t.src, t.tgt, t.sp_idx, t.sp_ptr, t.tp_idx, t.tp_ptr, t.tgt_patch_of_point,
                                        t.pair_tgt_patch, corr3d=t.corr3d, d_median_resolution=med,
                                        n_src_items=t.n_src_items, n_tgt_items=t.n_tgt_items, out=outs[i],
                                        **cfg.fine_kwargs())
    for s in streams:
        cur.wait_stream(s)


for S in (1, 2, 4, 8):
    streams = pipeline.make_streams(S, dev)
    for _ in range(2):
        run(streams)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run(streams)
    e1.record()
    torch.cuda.synchronize()
    print("streams=%d  %.3f ms per tile (%d pairs per tile)" % (S, e0.elapsed_time(e1) / 3 / NT, tiles[0].n_pairs), flush=True)
