"""GPU parity: the fused fine-matching stage (f4l_fine_matching) against the restatement of
src/coarse_to_fine_matching_base.py:3236-3457 (oracle/fine_matching.py) on a synthetic tile.
Integer decisions (selected correspondences, patch status, NN assignment) must be identical;
transforms and DVF rows within 1e-5 m (plus the f32 ulp of the coordinate for f32 rows)."""
import numpy as np
import pytest
import torch

from oracle import fine_matching as ofm
from oracle import knn as oknn

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _setup(n_pts, seed, patch_pts, dev, extra_2d=False):
    from fusion4landslide_b200 import synth
    d = synth.make_tile(n_pts, seed=seed, patch_pts=patch_pts)
    lab_s, ptr_s, idx_s = synth.patches_from_labels(d["label_src"])
    lab_t, ptr_t, idx_t = synth.patches_from_labels(d["label_tgt"])
    m, j = synth.pair_patches(lab_s, lab_t)
    spt_src = [idx_s[ptr_s[a]:ptr_s[a + 1]].numpy().astype(np.int64) for a in m.tolist()]
    spt_tgt = [idx_t[ptr_t[b]:ptr_t[b + 1]].numpy().astype(np.int64) for b in j.tolist()]
    sp_ptr = np.zeros(len(spt_src) + 1, np.int32)
    sp_ptr[1:] = np.cumsum([len(x) for x in spt_src])
    tp_ptr = np.zeros(len(spt_tgt) + 1, np.int32)
    tp_ptr[1:] = np.cumsum([len(x) for x in spt_tgt])
    tpo = -np.ones(d["tgt"].shape[0], np.int32)
    for b in range(lab_t.numel()):
        tpo[idx_t[ptr_t[b]:ptr_t[b + 1]].numpy()] = b
    g = dict(src=d["src"].to(dev), tgt=d["tgt"].to(dev), corr3d=d["corr3d"].to(dev),
             sp_idx=torch.from_numpy(np.concatenate(spt_src).astype(np.int32)).to(dev),
             sp_ptr=torch.from_numpy(sp_ptr).to(dev),
             tp_idx=torch.from_numpy(np.concatenate(spt_tgt).astype(np.int32)).to(dev),
             tp_ptr=torch.from_numpy(tp_ptr).to(dev), tpo=torch.from_numpy(tpo).to(dev),
             pair_tgt=j.to(torch.int32).to(dev))
    return d, spt_src, spt_tgt, g


def _ulp(x):
    return np.spacing(np.abs(x).astype(np.float32)).astype(np.float64)


# patch_pts 1400: more matched pairs than the warp kernel takes (CTA fit kernel) and target patches larger than
# the shared-memory staging area of the assign kernel (global-memory path)
@pytest.mark.parametrize("assign,tgt2src,patch_pts", [("assign_then_nn", True, 200), ("assign_all_src", False, 200),
                                                      ("assign_then_nn", True, 1400)])
def test_fine_matching_vs_oracle(cuda, assign, tgt2src, patch_pts):
    from fusion4landslide_b200 import ops
    d, spt_src, spt_tgt, g = _setup(80_000, 31, patch_pts, cuda)
    med = oknn.median_resolution(d["src"].numpy(), d["tgt"].numpy())
    prm = ofm.FineParams(mode="only_3d", assign_type=assign, output_tgt2src=tgt2src, median_max_resolution=med)
    o = ofm.fine_matching(d["src"].numpy(), d["tgt"].numpy(), d["corr3d"].numpy(), None, spt_src, spt_tgt, prm)
    r = ops.fine_matching(g["src"], g["tgt"], g["sp_idx"], g["sp_ptr"], g["tp_idx"], g["tp_ptr"], g["tpo"],
                          g["pair_tgt"], corr3d=g["corr3d"], assign_type=assign, output_tgt2src=tgt2src,
                          median_max_resolution=med)
    torch.cuda.synchronize()
    Q = len(spt_src)
    status = r.status.cpu().numpy()
    np.testing.assert_array_equal(r.K.cpu().numpy(), o["K"])
    np.testing.assert_array_equal(status, o["status"])
    assert (status == 0).sum() > Q // 2 and (status == 2).sum() >= 0
    np.testing.assert_allclose(r.dist_mean.cpu().numpy(), o["dist_mean"], atol=2e-5)
    np.testing.assert_allclose(r.ratio_inlier.cpu().numpy(), o["ratio_inlier"], atol=1e-3)
    T64 = r.T64.cpu().numpy()
    iters = r.iters.cpu().numpy()
    fitted = np.nonzero(status == 0)[0]
    same_path = (iters[fitted] == o["iters"][fitted]) & (np.abs(r.fitness.cpu().numpy()[fitted] - o["fitness"][fitted]) < 1e-12)
    n_flip = int((~same_path).sum())
    print("fine: %d pairs, %d fitted, %d ICP path flips" % (Q, fitted.size, n_flip))
    assert n_flip <= max(1, fitted.size // 200)
    src = d["src"].numpy()
    for q in fitted[same_path]:
        P = src[spt_src[q]].astype(np.float64)
        a = P @ T64[q, :3, :3].T + T64[q, :3, 3]
        b = P @ o["T64"][q, :3, :3].T + o["T64"][q, :3, 3]
        assert np.linalg.norm(a - b, axis=1).max() < TOL, q
    np.testing.assert_allclose(r.rmse.cpu().numpy()[fitted[same_path]], o["rmse"][fitted[same_path]], atol=1e-9)
    dense, sparse, t2s = r.rows()
    dense, sparse = dense.cpu().numpy(), sparse.cpu().numpy()
    od = ofm.stack(o["dense"])
    assert dense.shape == od.shape
    ok_rows = np.repeat(np.isin(np.arange(Q), fitted[same_path])[fitted], [len(spt_src[q]) for q in fitted])
    np.testing.assert_array_equal(dense[:, :3], od[:, :3])
    assert (np.abs(dense[ok_rows, 3:] - od[ok_rows, 3:]) <= TOL + 2 * _ulp(od[ok_rows, 3:])).all()
    if tgt2src:
        ot = ofm.stack(o["tgt2src"])
        t2s = t2s.cpu().numpy()
        assert t2s.shape == ot.shape
        np.testing.assert_array_equal(t2s[:, 3:], ot[:, 3:])
        ok_t = np.repeat(np.isin(np.arange(Q), fitted[same_path])[fitted], [len(spt_tgt[q]) for q in fitted])
        assert (np.abs(t2s[ok_t, :3] - ot[ok_t, :3]) <= TOL + 2 * _ulp(ot[ok_t, :3])).all()
    # sparse rows: compare pair by pair (row counts can differ only on flipped pairs / threshold ties)
    osp = [x for x in o["sparse"]]
    counts = [0 if x is None else x.shape[0] for x in osp]
    if n_flip == 0:
        want = ofm.stack(osp)
        if sparse.shape == want.shape:
            np.testing.assert_array_equal(sparse[:, :3], want[:, :3])
            if assign == "assign_then_nn":
                mism = (sparse[:, 3:] != want[:, 3:]).any(1).mean()
                assert mism < 2e-3, mism       # NN flips at fp64-rounding ties of the f32 query
            else:
                assert (np.abs(sparse[:, 3:] - want[:, 3:]) <= TOL + 2 * _ulp(want[:, 3:])).all()
        else:
            assert abs(sparse.shape[0] - want.shape[0]) <= max(4, want.shape[0] // 2000)
    assert r.counts.tolist()[3] == fitted.size
    assert sum(counts) > 0


def test_fine_matching_fusion_mode_and_device_median(cuda):
    """fusion = 3D rows then 2D rows; median resolution taken from the device scalar of A1."""
    from fusion4landslide_b200 import ops
    d, spt_src, spt_tgt, g = _setup(50_000, 32, 150, cuda)
    rng = np.random.default_rng(0)
    c2 = d["corr3d"].numpy().copy()
    drop = rng.random(c2.shape[0]) < 0.7
    c2[drop, 1] = -1
    med_dev = ops.median_resolution(g["src"], g["tgt"])
    med = oknn.median_resolution(d["src"].numpy(), d["tgt"].numpy())
    assert abs(med_dev.item() - med) < 2e-6 * med + 1e-7
    prm = ofm.FineParams(mode="fusion", median_max_resolution=float(np.float32(med_dev.item())))
    o = ofm.fine_matching(d["src"].numpy(), d["tgt"].numpy(), d["corr3d"].numpy(), c2, spt_src, spt_tgt, prm)
    r = ops.fine_matching(g["src"], g["tgt"], g["sp_idx"], g["sp_ptr"], g["tp_idx"], g["tp_ptr"], g["tpo"],
                          g["pair_tgt"], corr3d=g["corr3d"], corr2d=torch.from_numpy(c2).to(cuda), mode="fusion",
                          d_median_resolution=med_dev, want_fragile=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(r.K.cpu().numpy(), o["K"])
    np.testing.assert_array_equal(r.status.cpu().numpy(), o["status"])
    # the fused stage's ICP fragility flags (f4l_fine_buffers.icp_fragile): 0 for pairs that were not fitted, a 3-bit OR else;
    # a pair whose iteration count differs from the oracle's must carry one
    frag = r.icp_fragile.cpu().numpy()
    assert frag.shape == o["status"].shape and (frag[o["status"] != 0] == 0).all() and (frag < 8).all()
    flip = (o["status"] == 0) & (r.iters.cpu().numpy() != o["iters"])
    assert (frag[flip] != 0).all()
    dense, _, _ = r.rows()
    assert dense.shape[0] == ofm.stack(o["dense"]).shape[0]


def test_fine_matching_empty(cuda):
    from fusion4landslide_b200 import ops
    z = torch.zeros(0, dtype=torch.int32, device=cuda)
    p = torch.zeros(1, dtype=torch.int32, device=cuda)
    pts = torch.zeros(5, 3, device=cuda)
    r = ops.fine_matching(pts, pts, z, p, z, p, torch.zeros(5, dtype=torch.int32, device=cuda), z,
                          corr3d=torch.zeros(5, 2, dtype=torch.int64, device=cuda), n_src_items=0, n_tgt_items=0)
    assert r.counts.tolist() == [0, 0, 0, 0]


def test_prepare_tile_kernels_match_host_twin(cuda):
    """f4l_labels_to_csr / f4l_gather_pairs_csr (prepare_pts2spt_dict, base.py:1301-1351) against the torch
    restatement used for the CPU arm: identical tables."""
    from fusion4landslide_b200 import pipeline, synth
    d = synth.make_tile(50_000, seed=3, patch_pts=120)
    lab_s = d["label_src"].clone()
    lab_s[::97] = -5                               # a negative label and ...
    lab_s[5::1013] = 10**12 + 7                    # ... a huge one with <= 10 points
    host = pipeline.prepare_tile(d["src"], d["tgt"], lab_s, d["label_tgt"], d["corr3d"])
    dev = pipeline.prepare_tile(d["src"].to(cuda), d["tgt"].to(cuda), lab_s.to(cuda), d["label_tgt"].to(cuda),
                                d["corr3d"].to(cuda))
    torch.cuda.synchronize()
    assert (host.n_pairs, host.n_src_items, host.n_tgt_items) == (dev.n_pairs, dev.n_src_items, dev.n_tgt_items)
    for k in ("sp_ptr", "sp_idx", "tp_ptr", "tp_idx", "tgt_patch_of_point", "pair_tgt_patch"):
        assert torch.equal(getattr(host, k), getattr(dev, k).cpu()), k
    from fusion4landslide_b200 import ops
    e = ops.labels_to_csr(torch.zeros((0,), dtype=torch.int64, device=cuda))
    assert e[0].numel() == 0 and e[1].tolist() == [0] and e[2].numel() == 0


def test_batched_fit_over_tiles_equals_per_tile(cuda):
    """pipeline.displacement_field_tiles_batched (the rigid fits of all tiles in one queue-driven persistent launch,
    f4l_fine_fit_tiles) returns bit-identical results to the per-tile launches, with and without side streams, and
    again when the prepared calls are reused from the cache (second step)."""
    from fusion4landslide_b200 import pipeline, synth
    tiles = []
    for k, n in enumerate((9000, 30000, 14000)):
        d = synth.make_scene(n, seed=40 + k)
        tiles.append(pipeline.prepare_tile(d["src"].to(cuda), d["tgt"].to(cuda), d["label_src"].to(cuda), d["label_tgt"].to(cuda),
                                           d["corr3d"].to(cuda)))
    ref = pipeline.displacement_field_tiles(tiles)
    torch.cuda.synchronize()
    streams = pipeline.make_streams(2, cuda)
    sides = pipeline.make_streams(2, cuda)
    cache = {}
    for attempt in range(2):
        for kw in (dict(), dict(streams=streams, side_streams=sides, cache=cache)):
            got = pipeline.displacement_field_tiles_batched(tiles, **kw)
            torch.cuda.synchronize()
            for (r0, m0), (r1, m1) in zip(ref, got):
                assert torch.equal(m0, m1)
                for name in ("T", "T64", "status", "K", "fitness", "rmse", "iters", "ratio_inlier", "dist_mean", "counts"):
                    assert torch.equal(getattr(r0, name), getattr(r1, name)), name
                c = r0.counts.tolist()
                assert c[0] > 1000 and torch.equal(r0.dense[:c[0]], r1.dense[:c[0]])
                assert torch.equal(r0.sparse[:c[1]], r1.sparse[:c[1]])
    assert len(cache) == 3
