#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_exchange.py : every rank runs the hot path on its own tiles with the
fused exchange (exchange.PeerExchange) and checks its gathered field against an NCCL all-gather."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusion4landslide_b200 import pipeline, synth  # noqa: E402
from fusion4landslide_b200.exchange import PeerExchange  # noqa: E402
from fusion4landslide_b200.ops import FineResult  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tiles = []
    for t in range(rank, 2 * world, world):
        d = synth.make_tile(20_000 + 1000 * t, seed=100 + t, device=dev, patch_pts=200)
        tiles.append(pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"]))
    cap = torch.tensor([sum(t.n_src_items for t in tiles)], device=dev)
    dist.all_reduce(cap, op=dist.ReduceOp.MAX)
    cap_rows = int(cap[0])
    ex = PeerExchange(cap_rows, dev, n_buffers=2)
    for step in range(3):
        par = step % 2
        ex.field(par).fill_(float("nan"))
        torch.cuda.synchronize()
        dist.barrier()
        arena = ex.local_arena(par)
        ro, rows = 0, 0
        for t in tiles:
            r, _ = pipeline.displacement_field(t, peer_dense=ex.peer_ptrs(par, ro))   # default allocation ...
            n = int(r.counts[0])
            arena[ro:ro + n].copy_(r.dense[:n])                                       # ... copied into the local slice
            ro += t.n_src_items
            rows = ro
        cnt = torch.tensor([rows], device=dev)
        counts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(counts, cnt)                       # the closing collective = barrier for the pushed rows
        ref = torch.empty((world * cap_rows, 6), device=dev)
        dist.all_gather_into_tensor(ref, arena)
        ref = ref.view(world, cap_rows, 6)
        f = ex.field(par)
        for r_ in range(world):
            a, b = f[r_], ref[r_]
            m = ~torch.isnan(b[:, 0])
            assert m.any() and torch.equal(a[m], b[m]), "rank %d: slice %d differs at step %d" % (rank, r_, step)
    dist.barrier()
    ex.close()
    if rank == 0:
        print("exchange ok: %d ranks, %d rows per slice" % (world, cap_rows))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
