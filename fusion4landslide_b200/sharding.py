"""Multi-GPU plumbing of the hot path: tiles are independent (main_fusion.py:134-148 processes
them serially with no cross-tile state), so they are dealt to ranks by size (LPT) and each rank runs its
tiles with no data-path collective.  The ONE exchange step is the all-gather of the per-pair transforms and
the dense displacement field at the end (NCCL on GPUs; the same code runs on `gloo` for the CPU tests --
the collective is backend-agnostic, the kernels are not)."""
import heapq

import torch
import torch.distributed as dist


def lpt_assign(sizes, world):
    """Longest-processing-time-first: tile ids per rank, balanced by point count.  Deterministic: ties by tile id."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    out = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + int(sizes[i]), r))
    for lst in out:
        lst.sort()
    return out


class GatherPlan:
    """Fixed-shape buffers for the end-of-step exchange: every rank contributes `cap_rows` DVF rows,
    `cap_pairs` transforms and `cap_tiles` x 4 row counters (its arenas, padded to the max over ranks)."""

    def __init__(self, rows, pairs, tiles, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        caps = torch.tensor([rows, pairs, tiles], dtype=torch.int64, device=device)
        if self.world > 1:
            dist.all_reduce(caps, op=dist.ReduceOp.MAX, group=group)
        self.cap_rows, self.cap_pairs, self.cap_tiles = (int(x) for x in caps.tolist())
        self.dense = torch.zeros((self.cap_rows, 6), dtype=torch.float32, device=device)
        self.T = torch.zeros((self.cap_pairs, 4, 4), dtype=torch.float32, device=device)
        self.counts = torch.zeros((self.cap_tiles, 4), dtype=torch.int32, device=device)
        if self.world > 1:
            self.all_dense = torch.empty((self.world * self.cap_rows, 6), dtype=torch.float32, device=device)
            self.all_T = torch.empty((self.world * self.cap_pairs, 4, 4), dtype=torch.float32, device=device)
            self.all_counts = torch.empty((self.world * self.cap_tiles, 4), dtype=torch.int32, device=device)
        else:
            self.all_dense, self.all_T, self.all_counts = self.dense, self.T, self.counts

    def exchange(self):
        """all_gather_into_tensor of the three arenas (no-op for one rank)."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.all_T, self.T, group=self.group)
            dist.all_gather_into_tensor(self.all_dense, self.dense, group=self.group)
            dist.all_gather_into_tensor(self.all_counts, self.counts, group=self.group)
        return self.all_dense, self.all_T, self.all_counts

    def assemble(self, tile_rows, assignment):
        """Host-side view of the gathered field in GLOBAL tile order: list over tiles of (rows,6) slices.
        tile_rows[r] = list of arena row offsets of rank r's tiles (same order as assignment[r])."""
        counts = self.all_counts.reshape(self.world, self.cap_tiles, 4)
        out = {}
        for r, tiles in enumerate(assignment):
            for k, tile in enumerate(tiles):
                n = int(counts[r, k, 0])
                o = r * self.cap_rows + tile_rows[r][k]
                out[tile] = self.all_dense[o:o + n]
        return [out[t] for t in sorted(out)]
