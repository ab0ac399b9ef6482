"""CPU tests of the drop-in boundary: libf4l_b200.so loads, exports every symbol include/f4l_b200.h
declares, the ctypes table covers all of them, and the host layer fails loudly without a GPU
(no compute calls here)."""
import os
import re

import pytest
import torch

from fusion4landslide_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "f4l_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"F4L_API\s+[\w\s\*]+?\b(f4l_\w+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from fusion4landslide_b200 import build
        build.build()
    names = _declared()
    assert len(names) >= 25
    L = _lib.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    untyped = [n for n in names if n not in _lib.SIGNATURES]
    assert not untyped, untyped
    extra = [n for n in _lib.SIGNATURES if n not in names]
    assert not extra, extra
    assert L.f4l_abi_version() == 4
    assert L.f4l_launch_count() >= 0


def test_workspace_queries_are_host_functions():
    L = _lib.lib()
    assert L.f4l_knn_grid_workspace_bytes(1000, 1000) > 0
    assert L.f4l_desc_nn_workspace_bytes(1000, 2000, 64, 1) >= L.f4l_desc_nn_workspace_bytes(1000, 2000, 64, 0) > 0
    assert L.f4l_piecewise_icp_workspace_bytes(1000, 1000) > 0
    assert L.f4l_fine_matching_workspace_bytes(1000, 1000, 10, 0) > 0


def test_no_cpu_fallback():
    from fusion4landslide_b200 import ops
    x = torch.zeros((4, 3))
    with pytest.raises(_lib.F4LError):
        ops.knn_grid(x, x, 1)
    with pytest.raises(_lib.F4LError):
        _lib.ptr(x)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fusion4landslide_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_host_expand_sparse_restores_doubled_layout():
    """f4l_host_expand_sparse is a HOST function (no GPU): rows emitted once per pair -> the reference's
    [pair rows][pair rows] layout (base.py:3430,3436), any thread count, empty pairs included."""
    import numpy as np
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 40, 500).astype(np.int32)
    counts[[0, 7, 499]] = 0
    total = int(counts.sum())
    once = rng.normal(size=(total, 6)).astype(np.float32)
    want = []
    off = 0
    for c in counts:
        blk = once[off:off + c]
        want += [blk, blk]
        off += c
    want = np.concatenate(want)
    for nt in (1, 3, 16):
        out = torch.zeros((2 * total + 5, 6))
        n = ops.host_expand_sparse(torch.from_numpy(once), torch.from_numpy(counts), out, n_threads=nt)
        assert n == 2 * total
        np.testing.assert_array_equal(out[:n].numpy(), want)
        assert not out[n:].any()
    assert ops.host_expand_sparse(torch.zeros((0, 6)), torch.zeros(4, dtype=torch.int32), torch.zeros((0, 6))) == 0
