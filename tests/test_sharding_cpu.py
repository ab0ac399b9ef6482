"""CPU tests of the multi-GPU plumbing: LPT tile assignment and the end-of-step all-gather (gloo,
world size 2).  The exchanged payload is synthetic (the kernels need a GPU); what is covered is exactly the
code path bench.py / a multi-GPU driver runs around them."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fusion4landslide_b200 import sharding


def test_lpt_assign_balances_and_is_deterministic():
    sizes = [781250] * 64
    a = sharding.lpt_assign(sizes, 8)
    assert sorted(sum(a, [])) == list(range(64)) and all(len(x) == 8 for x in a)
    rng = np.random.default_rng(0)
    sizes = rng.integers(5000, 1_000_000, 37).tolist()
    for w in (1, 2, 3, 8):
        a = sharding.lpt_assign(sizes, w)
        assert sorted(sum(a, [])) == list(range(37))
        loads = [sum(sizes[i] for i in x) for x in a]
        assert max(loads) - min(loads) <= max(sizes)          # LPT bound
        assert a == sharding.lpt_assign(sizes, w)
    assert sharding.lpt_assign([], 4) == [[], [], [], []]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [1000, 400, 700, 900, 50]
        assign = sharding.lpt_assign(sizes, world)
        mine = assign[rank]
        rows = sum(sizes[t] for t in mine)
        plan = sharding.GatherPlan(rows, 3 * len(mine), len(mine), torch.device("cpu"))
        assert plan.cap_rows == max(sum(sizes[t] for t in a) for a in assign)
        offs, o = [], 0
        for k, t in enumerate(mine):
            n = sizes[t] - t                                   # pretend t rows were rejected
            plan.dense[o:o + n] = float(t) + torch.arange(n, dtype=torch.float32)[:, None] * 1e-3
            plan.counts[k, 0] = n
            plan.T[3 * k:3 * k + 3] = torch.eye(4) * (t + 1)
            offs.append(o)
            o += sizes[t]
        all_offs = [None] * world
        dist.all_gather_object(all_offs, offs)
        dense, T, counts = plan.exchange()
        field = plan.assemble(all_offs, assign)
        ok = len(field) == len(sizes)
        for t, rows_t in enumerate(field):
            ok &= rows_t.shape[0] == sizes[t] - t
            ok &= bool(torch.allclose(rows_t[:, 0], float(t) + torch.arange(sizes[t] - t, dtype=torch.float32) * 1e-3))
        for r, a in enumerate(assign):
            for k, t in enumerate(a):
                ok &= float(T[r * plan.cap_pairs + 3 * k, 0, 0]) == t + 1
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gather_plan_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]
