"""Multi-GPU exchange of the displacement vector field, fused into the producing kernel.

SURVEY 8(e): tiles are independent (main_fusion.py:134-148), so ranks own disjoint tiles and the only
exchange is the gather of the per-pair transforms and the dense DVF.  Instead of an all-gather after
the step, every rank allocates the WHOLE gathered field (world x cap_rows x 6 f32) as a peer-visible
buffer (f4l_peer_alloc -> CUDA IPC handle), maps the buffers of all other ranks (f4l_peer_open) and
hands the mapped pointers to f4l_fine_matching: the D5 kernel (k_apply_assign) stores each dense row
into the local field and into every peer's field over NVLink while the tile is being computed.  The
small NCCL all-gather of the transforms / row counts that follows the step doubles as the barrier that
makes all pushed rows visible (a rank enters it only after its own kernels, hence its stores, are
complete in stream order).

Buffers are double-buffered by step parity: a rank's pushes of step k+1 must not overwrite rows a
slower peer is still consuming from step k.
"""
import ctypes

import torch

from . import _lib


class _CudaBuffer:
    """Minimal __cuda_array_interface__ holder so torch can alias memory owned by libf4l_b200."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def alloc_peer_buffer(nbytes, device):
    """(device pointer, 64-byte IPC handle, uint8 tensor aliasing the buffer) on `device`."""
    L = _lib.lib()
    with torch.cuda.device(device):
        p = ctypes.c_void_p()
        h = ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        _lib.check(L.f4l_peer_alloc(int(nbytes), ctypes.byref(p), h), "f4l_peer_alloc")
        t = torch.as_tensor(_CudaBuffer(p.value, nbytes), device=torch.device(device))
    return p.value, h.raw, t


def open_peer_buffer(handle, device):
    L = _lib.lib()
    with torch.cuda.device(device):
        p = ctypes.c_void_p()
        _lib.check(L.f4l_peer_open(bytes(handle), ctypes.byref(p)), "f4l_peer_open")
    return p.value


class PeerExchange:
    """Gathered dense DVF of all ranks, written by the producers themselves.

    field(parity)            (world, cap_rows, 6) f32 tensor on this rank: slice r is rank r's arena
    local_arena(parity)      this rank's slice (what its own tiles write)
    peer_ptrs(parity, row)   device pointers of row `row` of THIS rank's slice in every other rank's field
    """

    ROW_BYTES = 24

    def __init__(self, cap_rows, device, group=None, n_buffers=2):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world - 1 > _lib.MAX_PEERS:
            raise _lib.F4LError("PeerExchange supports at most %d GPUs" % (_lib.MAX_PEERS + 1))
        self.device = torch.device(device)
        self.cap_rows = int(cap_rows)
        self.nbytes = self.world * self.cap_rows * self.ROW_BYTES
        self.local, self.remote = [], []          # per parity: (ptr, tensor) / {rank: mapped ptr}
        for _ in range(n_buffers):
            p, handle, t = alloc_peer_buffer(max(self.nbytes, 256), self.device)
            handles = [None] * self.world
            dist.all_gather_object(handles, handle, group=group)
            mapped = {r: open_peer_buffer(handles[r], self.device) for r in range(self.world) if r != self.rank}
            self.local.append((p, t))
            self.remote.append(mapped)
        dist.barrier(group=group)

    def field(self, parity=0):
        t = self.local[parity % len(self.local)][1]
        return t[:self.nbytes].view(torch.float32).view(self.world, self.cap_rows, 6)

    def local_arena(self, parity=0):
        return self.field(parity)[self.rank]

    def peer_ptrs(self, parity=0, row=0):
        off = (self.rank * self.cap_rows + int(row)) * self.ROW_BYTES
        return [base + off for _, base in sorted(self.remote[parity % len(self.remote)].items())]

    def close(self):
        """Collective: unmap the peers' buffers, then free our own."""
        L = _lib.lib()
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for mapped in self.remote:
                for p in mapped.values():
                    _lib.check(L.f4l_peer_close(p), "f4l_peer_close")
            self.remote = []
            self.dist.barrier(group=self.group)
            for p, _ in self.local:
                _lib.check(L.f4l_peer_free(p), "f4l_peer_free")
            self.local = []
