"""GPU parity: K-a grid kNN through the C ABI vs the exact fp64 oracle (scipy cKDTree) and the
reference's own sklearn calls (golden).  Indices must be identical except on rows flagged as
documented ties (adjacent squared distances within 1e-6 relative)."""
import os

import numpy as np
import pytest
import torch

from oracle import knn as oknn

pytestmark = pytest.mark.gpu


def _check(q, r, k, idx, d2, max_radius=None):
    oi, od2 = oknn.knn_exact(q, r, min(k, r.shape[0]))
    ties = oknn.tie_rows(q, r, min(k, r.shape[0]))
    idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
    kk = oi.shape[1]
    if max_radius is not None:
        inside = od2 < max_radius ** 2
        edge = np.abs(od2 - max_radius ** 2) <= 1e-6 * max_radius ** 2
        oi = np.where(inside, oi, -1)
        ties = ties | edge.any(1)
    same = (idx[:, :kk] == oi).all(1)
    assert (same | ties).all(), "index mismatch on %d non-tie rows" % int((~same & ~ties).sum())
    ok = idx[:, :kk] >= 0
    np.testing.assert_allclose(d2[:, :kk][ok], od2[ok], rtol=2e-6, atol=1e-10)
    if kk < k:
        assert (idx[:, kk:] == -1).all()
    return int(ties.sum())


def test_knn_golden_sklearn(cuda, golden_dir):
    from fusion4landslide_b200 import ops
    z = np.load(os.path.join(golden_dir, "knn_sklearn.npz"))
    a, b = torch.from_numpy(z["a"]).to(cuda), torch.from_numpy(z["b"]).to(cuda)
    idx, d2 = ops.knn_grid(a, a, 2)
    ties = oknn.tie_rows(z["a"], z["a"], 2)
    same = (idx.cpu().numpy() == z["self_i_a"]).all(1)
    assert (same | ties).all()
    np.testing.assert_allclose(np.sqrt(d2.cpu().numpy()), z["self_d_a"], rtol=1e-5, atol=1e-7)
    idx1, d21 = ops.knn_grid(a, b, 1)
    t1 = oknn.tie_rows(z["a"], z["b"], 1)
    assert ((idx1.cpu().numpy()[:, 0] == z["c2c_idx"][:, 0]) | t1).all()
    np.testing.assert_allclose(np.sqrt(d21.cpu().numpy()), z["c2c"], rtol=1e-5)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 8])
def test_knn_tile_self_and_cross(cuda, k):
    from fusion4landslide_b200 import ops, synth
    d = synth.make_tile(120_000, seed=11, device="cpu")
    src, tgt = d["src"], d["tgt"]
    i, dd = ops.knn_grid(src.to(cuda), src.to(cuda), k)
    _check(src.numpy(), src.numpy(), k, i, dd)
    i, dd = ops.knn_grid(src.to(cuda), tgt.to(cuda), k)
    _check(src.numpy(), tgt.numpy(), k, i, dd)


def test_knn_edge_cases(cuda):
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(12)
    # volumetric cloud, queries far outside the reference box, fewer refs than k, radius gate
    r = rng.uniform(-1, 1, size=(5000, 3)).astype(np.float32)
    q = np.vstack([rng.uniform(-3, 3, size=(3000, 3)), [[50, 50, 50], [-40, 0, 0]]]).astype(np.float32)
    i, dd = ops.knn_grid(torch.from_numpy(q).to(cuda), torch.from_numpy(r).to(cuda), 4)
    _check(q, r, 4, i, dd)
    i, dd = ops.knn_grid(torch.from_numpy(q).to(cuda), torch.from_numpy(r).to(cuda), 2, max_radius=0.3)
    _check(q, r, 2, i, dd, max_radius=0.3)
    r3 = r[:3]
    i, dd = ops.knn_grid(torch.from_numpy(q).to(cuda), torch.from_numpy(r3).to(cuda), 8)
    _check(q, r3, 8, i, dd)
    # duplicates: exact ties resolve to the lowest index
    rd = np.repeat(r[:100], 3, axis=0)
    i, dd = ops.knn_grid(torch.from_numpy(r[:100]).to(cuda), torch.from_numpy(rd).to(cuda), 3)
    assert (i.cpu().numpy() == (np.arange(100)[:, None] * 3 + np.arange(3))).all()
    # empty reference / empty query
    i, dd = ops.knn_grid(torch.from_numpy(q).to(cuda), torch.zeros(0, 3, device=cuda), 1)
    assert (i == -1).all()
    ops.knn_grid(torch.zeros(0, 3, device=cuda), torch.from_numpy(r).to(cuda), 1)
    # collinear / planar degenerate boxes
    line = np.zeros((2000, 3), np.float32)
    line[:, 0] = rng.uniform(0, 10, 2000)
    i, dd = ops.knn_grid(torch.from_numpy(line).to(cuda), torch.from_numpy(line).to(cuda), 2)
    _check(line, line, 2, i, dd)


def test_select_kth_and_median_resolution(cuda, golden_dir):
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(13)
    for n in (1, 2, 3, 1000, 65_537, 300_000):
        x = (rng.random(n) ** 2).astype(np.float32)
        if n > 10:
            x[rng.integers(0, n, 5)] = x[0]          # duplicates
        xs = np.sort(x)
        for k in sorted({0, (n - 1) // 2, n // 2, n - 1}):
            k2 = min(k + 1, n - 1)
            out = ops.select_kth(torch.from_numpy(x).to(cuda), k, k2).cpu().numpy()
            assert out[0] == xs[k] and out[1] == xs[k2], (n, k)
    # strided column of an (N,2) array, as the median-resolution path uses it
    y = rng.random((5000, 2)).astype(np.float32)
    out = ops.select_kth(torch.from_numpy(y).to(cuda), 2499, 2500, stride=2, offset=1).cpu().numpy()
    ys = np.sort(y[:, 1])
    assert out[0] == ys[2499] and out[1] == ys[2500]
    # A1 against the reference's sklearn call pattern (golden)
    z = np.load(os.path.join(golden_dir, "knn_sklearn.npz"))
    med = ops.median_resolution(torch.from_numpy(z["a"]).to(cuda), torch.from_numpy(z["b"]).to(cuda)).item()
    assert abs(med - z["median_resolution"][0]) < 1e-6 * z["median_resolution"][0] + 1e-8


def test_knn_tie_flags(cuda):
    """f4l_knn_grid_ties: the rows another exact search may order differently.  Against the oracle's tie rule on the same
    data: every oracle tie is flagged (the kernel's f32 gaps can flag a few more near the epsilon), exact duplicates and
    lattice points are flagged, well separated points are not."""
    from fusion4landslide_b200 import ops
    from oracle import knn as oknn
    rng = np.random.default_rng(5)
    r = rng.uniform(0, 20, (6000, 3)).astype(np.float32)
    r[100] = r[7]                                               # an exact duplicate: every query near it ties
    lattice = (np.stack(np.meshgrid(np.arange(12), np.arange(12), [0.0]), -1).reshape(-1, 3) * 0.5 + 30).astype(np.float32)
    r = np.vstack([r, lattice])
    q = np.vstack([r[:3000], lattice[:50] + np.float32(0.25)])   # lattice cell centres: four equidistant neighbours
    R, Qt = torch.from_numpy(r).to(cuda), torch.from_numpy(q).to(cuda)
    for k in (1, 2):
        tie = ops.knn_ties(Qt, R, k).cpu().numpy().astype(bool)
        o = oknn.tie_rows(q, r, k)
        assert (tie | ~o).all(), int((o & ~tie).sum())          # no oracle tie missed
        assert tie[-50:].all() and tie[7] and tie[100]
        assert (tie & ~o).sum() <= 5 and 0 < tie.mean() < 0.1
        idx, _ = ops.knn_grid(Qt, R, k)
        oi, _ = oknn.knn_exact(q, r, k)
        same = (idx.cpu().numpy() == oi).all(1)
        assert (same | tie).all()                               # the BASELINE gate: indices identical except flagged rows
