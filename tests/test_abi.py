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
    assert L.f4l_abi_version() == 2
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
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
