"""GPU parity of the composed path on a synthetic tile, through the host mirror of the reference stages
(BASELINE configs C2 / C3 / C4 at test scale): descriptor 1-NN -> magnitude gate + scatter (B2) ->
superpoint features -> mutual coarse matching under the coordinate gate (B3) [+ 2D-vote pairs and 2D-lifted
correspondences for the fusion mode, B4] -> fused fine matching (F2 F3 D2 E1 D5 A4), against the same
composition of the oracle's restatements.  Integer decisions exact, DVF rows within 1e-5 m."""
import numpy as np
import pytest
import torch

from oracle import desc_nn as odesc
from oracle import fine_matching as ofm
from oracle import knn as oknn

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _spt_tables(xyz, feat, labels, min_pts=10):
    """Per-superpoint centroid and mean descriptor (normalised), numpy fp32, patches in label order."""
    lab, inv, cnt = np.unique(labels, return_inverse=True, return_counts=True)
    keep = cnt > min_pts
    P = lab.size
    c = np.zeros((P, 3))
    f = np.zeros((P, feat.shape[1]))
    np.add.at(c, inv, xyz.astype(np.float64))
    np.add.at(f, inv, feat.astype(np.float64))
    c /= cnt[:, None]
    f /= np.maximum(np.linalg.norm(f, axis=1, keepdims=True), 1e-12)
    spt = [np.nonzero(inv == p)[0] for p in np.nonzero(keep)[0]]
    return lab[keep], c[keep].astype(np.float32), f[keep].astype(np.float32), spt


@pytest.mark.parametrize("mode,D", [("only_3d", 64), ("fusion", 32)])
def test_coarse_to_fine_pipeline_vs_oracle(cuda, mode, D):
    from fusion4landslide_b200 import coarse_to_fine as c2f
    from fusion4landslide_b200 import pipeline, synth
    d = synth.make_tile(40_000, seed=17, patch_pts=220, desc_dim=D)
    src, tgt = d["src"].numpy(), d["tgt"].numpy()
    fs, ft = d["src_feat"].numpy(), d["tgt_feat"].numpy()
    ls, lt = d["label_src"].numpy(), d["label_tgt"].numpy()
    n = src.shape[0]
    max_mag = 5.0
    # ---- oracle composition -----------------------------------------------------------------
    ident = np.arange(n)
    C3, labels_o, keep_o = odesc.global_matches_from_3d(fs, ft, src, tgt, ident, ident, n, max_mag)
    lab_s, cs, fsp, spt_s = _spt_tables(src, fs, ls)
    lab_t, ct, ftp, spt_t = _spt_tables(tgt, ft, lt)
    m_o, j_o = odesc.coarse_matching_3d(cs, fsp, ct, ftp, max_mag, "nn_mutual")
    corr2d = None
    pairs_o = (m_o, j_o)
    if mode == "fusion":
        rng = np.random.default_rng(3)
        corr2d = -np.ones((n, 2), np.int64)
        corr2d[:, 0] = ident
        has = rng.random(n) < 0.08
        corr2d[has, 1] = d["gt_tgt_of_src"].numpy()[has]
        wrong = has & (rng.random(n) < 0.05)
        corr2d[wrong, 1] = rng.integers(0, n, wrong.sum())
        m2, j2, tie2 = odesc.coarse_matching_2d_vote(corr2d, lt, spt_s, lab_t)
        pairs_o = (np.concatenate([m2, m_o]), np.concatenate([j2, j_o]))       # 2D pairs then 3D pairs (base.py:3139-3146)
    med = oknn.median_resolution(src, tgt)
    prm = ofm.FineParams(mode=mode, median_max_resolution=float(np.float32(med)))
    o = ofm.fine_matching(src, tgt, C3, corr2d, [spt_s[a] for a in pairs_o[0]], [spt_t[b] for b in pairs_o[1]], prm)
    # ---- B200 path through the host mirror ------------------------------------------------------
    T = lambda x, dt=None: torch.from_numpy(np.ascontiguousarray(x)).to(cuda) if dt is None else \
        torch.from_numpy(np.ascontiguousarray(x)).to(cuda, dt)
    ident_t = torch.arange(n, device=cuda)
    C3g, labels = c2f.global_matches_from_3d(T(fs), T(ft), T(src), T(tgt), ident_t, ident_t, n, max_mag, algo="tensor")
    np.testing.assert_array_equal(labels.cpu().numpy(), labels_o)
    np.testing.assert_array_equal(C3g.cpu().numpy(), C3)
    m, j = c2f.coarse_matching_3d(cs, fsp, ct, ftp, max_mag, "nn_mutual")
    np.testing.assert_array_equal(m.cpu().numpy(), m_o)
    np.testing.assert_array_equal(j.cpu().numpy(), j_o)
    pairs = (m, j)
    c2d = None
    if mode == "fusion":
        c2d = T(corr2d)
        plab, pptr, pidx = synth.patches_from_labels(T(ls))
        mm, jj, tie = c2f.coarse_matching_2d(c2d, pidx, pptr, T(lt), T(lab_t))
        np.testing.assert_array_equal(mm.cpu().numpy(), m2)
        np.testing.assert_array_equal(jj.cpu().numpy(), j2)
        pairs = (torch.cat([mm, m]), torch.cat([jj, j]))
    r, tile = c2f.fine_matching_with_different_types(T(src), T(tgt), T(ls), T(lt), C3g, c2d, pairs=pairs,
                                                      config=pipeline.FineConfig(mode=mode))
    torch.cuda.synchronize()
    dense, sparse, _ = r.rows()
    np.testing.assert_array_equal(r.K.cpu().numpy(), o["K"])
    np.testing.assert_array_equal(r.status.cpu().numpy(), o["status"])
    od = ofm.stack(o["dense"])
    dn = dense.cpu().numpy()
    assert dn.shape == od.shape and dn.shape[0] > 0.5 * n
    same = r.iters.cpu().numpy() == o["iters"]
    assert same.mean() > 0.99
    sizes = [len(spt_s[a]) for q, a in enumerate(pairs_o[0]) if o["status"][q] == 0]
    rows_ok = np.repeat(same[o["status"] == 0], sizes)
    tol = TOL + 2 * np.spacing(np.float32(np.abs(od).max()))
    assert np.abs(dn[rows_ok] - od[rows_ok]).max() < tol
    osp = ofm.stack(o["sparse"])
    assert abs(sparse.shape[0] - osp.shape[0]) <= 0.002 * osp.shape[0] + 2
