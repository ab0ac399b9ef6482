"""GPU parity: K-g, the reference's Piecewise_ICP (src/piecewise_icp.py:89-216) through the C ABI against
the fp64 restatement oracle/piecewise.py (octree part parity-unpinned, see the oracle header).
Bar: cell tables and row order exact; centroids / displaced points within 1e-9 m (fp64 path; the
north-star tolerance is 1e-5 m)."""
import numpy as np
import pytest
import torch

from oracle import piecewise as opw

pytestmark = pytest.mark.gpu


def _scene(n, seed, shift=(0.3, -0.2, 0.1), L=40.0):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0, L, (n, 2))
    z = 2.0 * np.sin(xy[:, 0] * 0.3) + 1.5 * np.cos(xy[:, 1] * 0.2)
    src = np.column_stack([xy, z])
    xy2 = rng.uniform(0, L, (n + 137, 2))
    z2 = 2.0 * np.sin(xy2[:, 0] * 0.3) + 1.5 * np.cos(xy2[:, 1] * 0.2)
    tgt = np.column_stack([xy2, z2])
    moving = (tgt[:, 0] > L / 2) & (tgt[:, 1] > L / 2)
    tgt[moving] += np.asarray(shift)
    return src, tgt


@pytest.mark.parametrize("n,smax,min_pts,seed", [(60000, 5.0, 10, 0), (200000, 2.5, 10, 1), (30000, 5.0, 40, 2)])
def test_piecewise_icp_matches_oracle(cuda, n, smax, min_pts, seed):
    from fusion4landslide_b200 import ops
    src, tgt = _scene(n, seed)
    o = opw.piecewise_icp(src, tgt, smax, min_pts)
    dvfs, mag, counts, thr, cs, ct, nn = ops.piecewise_icp(torch.from_numpy(src).to(cuda), torch.from_numpy(tgt).to(cuda),
                                                           smax, min_pts, want_tables=True)
    torch.cuda.synchronize()
    c = counts.cpu().numpy()
    assert c[4] == o["depth"]
    assert c[2] == o["centroids_src"].shape[0] and c[3] == o["centroids_tgt"].shape[0]
    np.testing.assert_allclose(cs.cpu().numpy()[:c[2]], o["centroids_src"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(ct.cpu().numpy()[:c[3]], o["centroids_tgt"], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(nn.cpu().numpy()[:c[2]], o["nn"])
    assert abs(thr.item() - o["thr"]) < 1e-12
    assert c[5] == int((~o["stable"]).sum()) and c[5] > 0
    assert c[0] == o["dvfs"].shape[0] and c[1] == o["n_stable_pts"]
    np.testing.assert_allclose(dvfs.cpu().numpy()[:c[0]], o["dvfs"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(mag.cpu().numpy()[:c[0]], o["dvfms"][:, 3], rtol=0, atol=1e-9)
