"""GPU parity: voxel subsampling (Open3D voxel_down_sample as used by base.py:1012-1057) through the C ABI against
oracle/voxel.py (parity unpinned: Open3D absent; published algorithm).  Bar: voxel membership and centroids
bit-exact (fp64 sums in point order), rows in ascending voxel order."""
import numpy as np
import pytest
import torch

from oracle import knn as oknn
from oracle import voxel as ovox

pytestmark = pytest.mark.gpu


def _check(cuda, pts, voxel):
    from fusion4landslide_b200 import ops
    cent, vop = ops.voxel_downsample(torch.from_numpy(pts).to(cuda), voxel, want_map=True)
    oc, oinv = ovox.voxel_down_sample(pts, voxel)
    assert cent.shape[0] == oc.shape[0]
    np.testing.assert_array_equal(vop.cpu().numpy(), oinv)
    np.testing.assert_array_equal(cent.cpu().numpy(), oc)
    return cent


def test_voxel_downsample_tile(cuda):
    from fusion4landslide_b200 import synth
    d = synth.make_tile(300_000, seed=51, device="cpu")
    pts = d["src"].double().numpy() + np.array([2_600_000.0, 1_200_000.0, 500.0])     # national-grid sized coordinates
    for voxel in (0.05, 0.1, 0.37):
        c = _check(cuda, pts, voxel)
        assert 0 < c.shape[0] <= pts.shape[0]


def test_voxel_downsample_edge_cases(cuda):
    from fusion4landslide_b200 import ops, _lib
    rng = np.random.default_rng(3)
    _check(cuda, rng.uniform(-5, 5, (1, 3)), 0.3)                                       # single point
    p = rng.uniform(-5, 5, (2000, 3))
    _check(cuda, np.vstack([p, p[:500], p[:100]]), 0.25)                               # duplicates
    _check(cuda, np.round(rng.uniform(-3, 3, (5000, 3)) / 0.5) * 0.5, 0.5)             # points ON voxel borders
    _check(cuda, rng.uniform(0, 1, (3000, 3)) * np.array([100.0, 0.0, 1.0]), 0.2)      # degenerate (flat) axis
    _check(cuda, rng.uniform(-1, 1, (4000, 3)), 50.0)                                  # one voxel holds everything
    out = ops.voxel_downsample(torch.zeros((0, 3), dtype=torch.float64, device=cuda), 0.1)
    assert out.shape == (0, 3)
    with pytest.raises(_lib.F4LError):
        ops.voxel_downsample(torch.zeros((4, 3), dtype=torch.float64, device=cuda), 0.0)
    with pytest.raises(_lib.F4LError):
        ops.voxel_downsample(torch.zeros((4, 3), dtype=torch.float32, device=cuda), 0.1)
    far = np.array([[0.0, 0.0, 0.0], [1e9, 0.0, 0.0]])
    with pytest.raises(_lib.F4LError):
        ops.voxel_downsample(torch.from_numpy(far).to(cuda), 0.1)                      # > 2^21 voxels along x


def test_voxel_subsampling_mirror(cuda):
    """base.py:1012-1057 composed: adaptive voxel size, voxel means, nearest raw point maps."""
    from fusion4landslide_b200 import coarse_to_fine as c2f
    from fusion4landslide_b200 import synth
    d = synth.make_tile(60_000, seed=52, device="cpu")
    src, tgt = d["src"], d["tgt"]
    r = c2f.voxel_subsampling(src.to(cuda), tgt.to(cuda))
    med = oknn.median_resolution(src.numpy(), tgt.numpy())
    assert abs(r["voxel_size"] - med) <= 1e-6 * med
    for name, p in (("src", src), ("tgt", tgt)):
        oc, _ = ovox.voxel_down_sample(p.double().numpy(), r["voxel_size"])
        sub = r[name + "_pts_sub"].cpu().numpy()
        np.testing.assert_array_equal(sub, oc.astype(np.float32))
        oi, _ = oknn.knn_exact(sub, p.numpy(), 1)
        ties = oknn.tie_rows(sub, p.numpy(), 1)
        v2p = r["idx_voxel2pts_" + name].cpu().numpy()
        assert ((v2p == oi[:, 0]) | ties).all()
        p2v = r["idx_pts2voxel_" + name].cpu().numpy()
        assert p2v.shape[0] == p.shape[0] and (p2v[v2p] >= 0).all() and (p2v >= -1).all()
        want = -np.ones(p.shape[0], np.int64)
        np.maximum.at(want, v2p, np.arange(v2p.size))              # duplicates: the largest voxel index wins
        np.testing.assert_array_equal(p2v, want)
