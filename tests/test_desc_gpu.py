"""GPU parity: K-b descriptor-space nearest neighbour (tensor-core path and fp64 brute-force path),
coarse mutual matching under the xyz gate, and the global-match scatter, through the C ABI.

Bar (north_star): indices bit-exact except documented ties -- rows whose best and second-best
squared L2 distance differ by <= EPS_DESC_ABS = 1e-6 in the fp64 oracle.  The reference's own result
(torch.cdist + min on fp32, tests/golden/desc_cdist.npz) is compared under the same rule."""
import os

import numpy as np
import pytest
import torch

from oracle import desc_nn as odesc

pytestmark = pytest.mark.gpu
EPS = odesc.EPS_DESC_ABS


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def _check(idx, d2, a, b, what):
    oi, od, od2 = odesc.desc_nn(a, b, return_second=True)
    idx = idx.cpu().numpy().astype(np.int64)
    d2 = d2.cpu().numpy()
    bad = idx != oi
    ties = (od2 - od) <= EPS
    assert not (bad & ~ties).any(), (what, int((bad & ~ties).sum()), "mismatches outside documented ties")
    # a mismatch inside a tie must still be a minimiser within the tie epsilon
    if bad.any():
        rows = np.nonzero(bad)[0]
        dd = ((a[rows].astype(np.float64) - b[idx[rows]].astype(np.float64)) ** 2).sum(1)
        assert (dd - od[rows] <= EPS).all(), what
    np.testing.assert_allclose(d2, od, rtol=1e-5, atol=1e-6, err_msg=what)
    return int(bad.sum()), int(ties.sum())


@pytest.mark.parametrize("algo", ["exact", "tensor"])
@pytest.mark.parametrize("D", [32, 64])
def test_desc_nn_golden(cuda, golden_dir, D, algo):
    """Inputs and outputs of the reference's exact branch (base.py:2783-2815)."""
    from fusion4landslide_b200 import ops
    z = np.load(os.path.join(golden_dir, "desc_cdist.npz"))
    a, b = z["D%d_a" % D], z["D%d_b" % D]
    idx, d2 = ops.desc_nn(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda), algo=algo)
    torch.cuda.synchronize()
    _check(idx, d2, a, b, "golden D=%d %s" % (D, algo))
    # the reference's own labels (fp32 cdist): identical outside ties
    oi, od, od2 = odesc.desc_nn(a, b, return_second=True)
    ref = z["D%d_labels" % D].astype(np.int64)
    diff = idx.cpu().numpy() != ref
    assert not (diff & ((od2 - od) > 1e-5)).any()
    np.testing.assert_allclose(np.sqrt(d2.cpu().numpy()), z["D%d_dist" % D], atol=2e-4)


@pytest.mark.parametrize("D", [32, 64])
def test_desc_nn_tensor_path_shapes(cuda, D):
    """Tensor-core path on ragged sizes (N, M not multiples of the 256/128 tiles; several row blocks per
    CTA, > 4 pipeline stages), matched pairs + outliers like the synthetic benchmark descriptors."""
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(11 + D)
    for N, M in [(1, 1), (5, 700), (300, 129), (2049, 3001), (40000, 9000)]:
        b = _unit(rng, M, D)
        a = _unit(rng, N, D)
        k = min(N, M) // 2
        if k:
            noisy = b[:k] + 0.15 * rng.standard_normal((k, D)).astype(np.float32)
            a[:k] = noisy / np.linalg.norm(noisy, axis=1, keepdims=True)
        idx, d2 = ops.desc_nn(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda), algo="tensor")
        torch.cuda.synchronize()
        _check(idx, d2, a, b, "tensor N=%d M=%d D=%d" % (N, M, D))


def test_desc_nn_tensor_adversarial(cuda):
    """Duplicated reference rows (exact ties -> lowest index), a zero query (everything ties -> overflow ->
    fp64 brute force), unnormalised magnitudes (power-of-two rescale), near-duplicates below the fp16
    resolution (must be separated by the fp64 re-rank)."""
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(3)
    D, N, M = 64, 1500, 2600
    b = _unit(rng, M, D)
    b[100] = b[7]
    b[2000] = b[7]                                # three identical rows: 7 must win
    b[1234] = b[55] * (1 + 3e-4)                  # differs from row 55 below fp16 resolution
    a = _unit(rng, N, D)
    a[0] = b[7]
    a[1] = b[55]
    a[2] = b[1234]
    a[3] = 0.0                                    # all references at the same distance (unit norm) up to rounding
    a[4:40] = b[7] + 1e-3 * rng.standard_normal((36, D)).astype(np.float32)
    for scale in (1.0, 37.5, 1e-3):
        A, B = (a * scale).astype(np.float32), (b * scale).astype(np.float32)
        idx, d2 = ops.desc_nn(torch.from_numpy(A).to(cuda), torch.from_numpy(B).to(cuda), algo="tensor")
        torch.cuda.synchronize()
        i = idx.cpu().numpy()
        oi, od = odesc.desc_nn(A, B)
        assert i[0] == 7 and i[1] == 55 and i[2] == 1234
        assert (i[4:40] == 7).all()
        # fp64 argmin everywhere, evaluated by direct differences on the same f32 inputs; where the choice
        # differs from the oracle's it must be an equally good minimiser with a LOWER index (identical rows)
        # or a strictly better one (the oracle's GEMM-form fp64 distance carries ~1e-16 rounding)
        dd = ((A.astype(np.float64) - B[i].astype(np.float64)) ** 2).sum(1)
        do = ((A.astype(np.float64) - B[oi].astype(np.float64)) ** 2).sum(1)
        assert (dd <= do).all()
        neq = i != oi
        assert ((dd[neq] < do[neq]) | (i[neq] < oi[neq])).all()
        assert neq.sum() <= 64


@pytest.mark.parametrize("D", [32, 64])
def test_desc_nn_tensor_mixed_norms_use_the_bias_columns(cuda, D):
    """Reference rows of very different norms: 1/2||b||^2 is not a constant, the kernel variant WITH the bias
    K-columns must run (k_desc_choose); rows of (almost) equal norm but a spread just below the switch point run the
    variant without them with the spread added to the margin.  Both exact."""
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(21 + D)
    N, M = 3000, 5300
    for spread in (0.5, 1.5e-4, 0.0):
        b = _unit(rng, M, D) * (1.0 + spread * rng.uniform(-1, 1, size=(M, 1))).astype(np.float32)
        a = _unit(rng, N, D)
        a[:1500] = b[:1500] + 0.1 * rng.standard_normal((1500, D)).astype(np.float32)
        idx, d2 = ops.desc_nn(torch.from_numpy(a).to(cuda), torch.from_numpy(b.astype(np.float32)).to(cuda), algo="tensor")
        torch.cuda.synchronize()
        _check(idx, d2, a, b.astype(np.float32), "spread %g D=%d" % (spread, D))


def test_desc_nn_both_dirs_and_auto(cuda):
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(5)
    a, b = _unit(rng, 1800, 32), _unit(rng, 2100, 32)
    A, B = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
    r1 = ops.desc_nn(A, B, both_dirs=True, algo="tensor")
    r2 = ops.desc_nn(A, B, both_dirs=True, algo="exact")
    r3 = ops.desc_nn(A, B, both_dirs=True)
    torch.cuda.synchronize()
    _check(r1[0], r1[1], a, b, "rows")
    _check(r1[2], r1[3], b, a, "cols")
    for x, y, w in zip(r1, r2, r3):
        assert torch.equal(x, y) or x.dtype == torch.float32
        assert torch.equal(y, w)


def test_coarse_matching_golden(cuda, golden_dir):
    """base.py:2966-2995: feature NN under the coordinate gate, mutual test (B3)."""
    from fusion4landslide_b200 import ops
    z = np.load(os.path.join(golden_dir, "desc_cdist.npz"))
    cs, ct, fs, ft = z["coarse_cs"], z["coarse_ct"], z["coarse_fs"], z["coarse_ft"]
    mm = float(z["coarse_max_mag"][0])
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(cuda)
    ri, rd, ci, cd = ops.desc_nn(T(fs), T(ft), a_xyz=T(cs), b_xyz=T(ct), max_mag=mm, both_dirs=True)
    torch.cuda.synchronize()
    ri, ci = ri.cpu().numpy(), ci.cpu().numpy()
    in_mag = ri >= 0
    np.testing.assert_array_equal(in_mag, z["coarse_in_mag"].astype(bool))
    np.testing.assert_array_equal(ri[in_mag], z["coarse_j"][in_mag])
    mutual = in_mag & (ci[np.maximum(ri, 0)] == np.arange(ri.shape[0]))
    np.testing.assert_array_equal(mutual, z["coarse_mutual"].astype(bool) & in_mag)
    m, j = odesc.coarse_matching_3d(cs, fs, ct, ft, mm)
    np.testing.assert_array_equal(np.nonzero(mutual)[0], m)
    np.testing.assert_array_equal(ri[mutual], j)


def test_scatter_global_matches(cuda):
    """base.py:2872-2889 incl. duplicate raw targets (largest voxel index wins)."""
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(9)
    n_sub, m_sub, n_raw, m_raw, D = 5000, 5200, 9000, 9500, 32
    fs, ft = _unit(rng, n_sub, D), _unit(rng, m_sub, D)
    src_sub = rng.uniform(0, 20, (n_sub, 3)).astype(np.float32)
    tgt_sub = rng.uniform(0, 20, (m_sub, 3)).astype(np.float32)
    v2p_s = rng.integers(0, n_raw, n_sub)          # with duplicates
    v2p_t = rng.integers(0, m_raw, m_sub)
    C, labels, keep = odesc.global_matches_from_3d(fs, ft, src_sub, tgt_sub, v2p_s, v2p_t, n_raw, 8.0)
    T = lambda x: torch.from_numpy(x).to(cuda)
    lab, _ = ops.desc_nn(T(fs), T(ft))
    np.testing.assert_array_equal(lab.cpu().numpy(), labels)
    out = ops.scatter_global_matches(lab, T(src_sub), T(tgt_sub), T(v2p_s), T(v2p_t), 8.0, n_raw)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), C)


def test_desc_nn_tie_flags(cuda):
    """f4l_desc_nn_ex: rows with another reference row within eps of the minimum (duplicates, a zero query) are flagged on
    both kernel paths; generic rows are not.  The BASELINE gate: labels equal the golden except flagged rows."""
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(8)
    D, N, M = 32, 2500, 4000
    b = _unit(rng, M, D)
    b[3000] = b[11]                                  # duplicate reference rows
    a = _unit(rng, N, D)
    a[0] = b[11] + 1e-3 * rng.standard_normal(D).astype(np.float32)
    a[1] = 0.0                                       # equidistant from every unit row up to rounding
    for algo in ("tensor", "exact"):
        idx, d2, tie = ops.desc_nn(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda), algo=algo, tie_eps=1e-6)
        torch.cuda.synchronize()
        t = tie.cpu().numpy().astype(bool)
        assert t[0] and idx[0].item() == 11          # the lower index of the duplicates, flagged
        assert t[1]
        oi, od, od2 = odesc.desc_nn(a, b, return_second=True)
        o_tie = (od2 - od) <= 1e-6
        assert (t | ~o_tie).all() and (t & ~o_tie).sum() <= 3
        assert ((idx.cpu().numpy() == oi) | t).all()
        assert t.mean() < 0.01
