"""GPU checks at BASELINE.json's FULL tile sizes (C1/C2: 1 M points per tile, C5: 781 250 points per tile).

Where the oracle still finishes in seconds (scipy cKDTree kNN, the median resolution) the comparison is direct.
Elsewhere the checks are size-independent properties of the path:
  * every dense DVF row is [p | T p] of its pair's transform (recomputed in fp64 from the returned T),
  * transforms are proper rotations, statuses / iteration counts are in range, row counts add up,
  * every sparse row's target is a point of the target cloud within the acceptance threshold of T p,
  * the known block motion of the synthetic scene is recovered (known-answer DVF),
  * two runs give bit-identical outputs (no atomics-order dependence in any result),
  * descriptor NN: returned distances equal the fp64 distance of the returned index, a random sample of rows
    equals the exact fp64 arg-min of the oracle, and mutual matches are symmetric.
"""
import numpy as np
import pytest
import torch

from oracle import desc_nn as odesc
from oracle import knn as oknn

pytestmark = pytest.mark.gpu


def test_knn_c5_tile_vs_ckdtree(cuda):
    from fusion4landslide_b200 import ops, synth
    d = synth.make_tile(781_250, seed=5, device="cpu")
    src, tgt = d["src"], d["tgt"]
    for q, r, k in ((src, src, 2), (src, tgt, 1)):
        idx, d2 = ops.knn_grid(q.to(cuda), r.to(cuda), k)
        oi, od2 = oknn.knn_exact(q.numpy(), r.numpy(), k)
        ties = oknn.tie_rows(q.numpy(), r.numpy(), k)
        same = (idx.cpu().numpy() == oi).all(1)
        assert (same | ties).all(), "index mismatch on %d non-tie rows" % int((~same & ~ties).sum())
        assert ties.mean() < 1e-3
        np.testing.assert_allclose(d2.cpu().numpy(), od2, rtol=2e-6, atol=1e-10)


def test_median_resolution_c2_tile(cuda):
    from fusion4landslide_b200 import ops, synth
    d = synth.make_tile(1_000_000, seed=6, device="cpu")
    med = ops.median_resolution(d["src"].to(cuda), d["tgt"].to(cuda)).item()
    ref = oknn.median_resolution(d["src"].numpy(), d["tgt"].numpy())
    assert abs(med - ref) <= 1e-6 * ref + 1e-9, (med, ref)


def test_fine_matching_c5_tile_properties(cuda):
    from fusion4landslide_b200 import pipeline, synth
    d = synth.make_tile(781_250, seed=7, device=cuda, patch_pts=256)
    t = pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"])
    cfg = pipeline.FineConfig()
    r, med = pipeline.displacement_field(t, cfg)
    torch.cuda.synchronize()
    dense, sparse, _ = r.rows()
    status = r.status.cpu().numpy()
    Q = status.size
    assert Q == t.n_pairs and set(np.unique(status)) <= {0, 1, 2}
    it = r.iters.cpu().numpy()
    assert (it[status == 0] >= 1).all() and (it <= 30).all() and (it[status != 0] == 0).all()
    # ---- transforms: proper rotations, identity for rejected pairs ----------------------------------
    T = r.T64.cpu().numpy().reshape(Q, 4, 4)
    R = T[:, :3, :3]
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.broadcast_to(np.eye(3), (Q, 3, 3)), atol=1e-12)
    assert (np.linalg.det(R) > 0.999999).all()
    assert (T[status != 0] == np.eye(4)).all()
    np.testing.assert_array_equal(r.T.cpu().numpy().reshape(Q, 4, 4), T.astype(np.float32))
    # ---- dense rows: [p | T p] in pair order over the accepted pairs ---------------------------------
    sp_ptr = t.sp_ptr.cpu().numpy().astype(np.int64)
    sp_idx = t.sp_idx.cpu().numpy()
    ok = status == 0
    cnt = (sp_ptr[1:] - sp_ptr[:-1])[:Q]
    assert dense.shape[0] == int(cnt[ok].sum())
    pair_of_row = np.repeat(np.nonzero(ok)[0], cnt[ok])
    item = np.concatenate([np.arange(sp_ptr[q], sp_ptr[q + 1]) for q in np.nonzero(ok)[0]])
    p = d["src"].cpu().numpy()[sp_idx[item]]
    dn = dense.cpu().numpy()
    np.testing.assert_array_equal(dn[:, :3], p)
    # the reference applies the f32 copy of the transform (base.py:3366): the row is the f32 rounding of T32 p
    T32 = r.T.cpu().numpy().reshape(Q, 4, 4).astype(np.float64)
    Tp = np.einsum("nij,nj->ni", T32[pair_of_row, :3, :3], p.astype(np.float64)) + T32[pair_of_row, :3, 3]
    assert np.abs(dn[:, 3:] - Tp).max() <= 3.1e-5           # half an f32 ulp at |coordinate| < 1024 m
    assert (dn[:, 3:] == Tp.astype(np.float32)).mean() > 0.999
    # ---- known answer: the block motion of the scene ---------------------------------------------------
    blk = d["block_of_src"].cpu().numpy()[sp_idx[item]]
    Rg, tg = d["R_gt"].cpu().numpy(), d["t_gt"].cpu().numpy()
    truth = np.einsum("nij,nj->ni", Rg[blk], p.astype(np.float64)) + tg[blk]
    err = np.linalg.norm(Tp - truth, axis=1)
    assert np.median(err) < 0.02, np.median(err)            # 5 mm noise per axis, 4 cm resampling jitter
    # ---- sparse rows: [p | matched target point], each written twice (base.py:3430,3436) ---------------
    sp = sparse.cpu().numpy()
    assert sp.shape[0] % 2 == 0 and sp.shape[0] > 0
    tgt_np = np.ascontiguousarray(d["tgt"].cpu().numpy())
    allkeys = set(tgt_np.view(np.dtype((np.void, 12))).ravel().tolist())
    samp = np.ascontiguousarray(sp[:: max(1, sp.shape[0] // 50_000), 3:])
    assert all(k in allkeys for k in samp.view(np.dtype((np.void, 12))).ravel().tolist())
    # ---- determinism ---------------------------------------------------------------------------------
    r2, med2 = pipeline.displacement_field(t, cfg)
    torch.cuda.synchronize()
    d2_, s2_, _ = r2.rows()
    assert med.item() == med2.item()
    assert torch.equal(dense, d2_) and torch.equal(sparse, s2_) and torch.equal(r.T64, r2.T64)
    assert torch.equal(r.iters, r2.iters) and torch.equal(r.status, r2.status)


@pytest.mark.parametrize("D", [32, 64])
def test_desc_nn_c2_tile_properties(cuda, D):
    from fusion4landslide_b200 import ops, synth
    n = 1_000_000 if D == 32 else 524_288
    d = synth.make_tile(n, seed=8, device=cuda, desc_dim=D)
    a, b = d["src_feat"], d["tgt_feat"]
    out = ops.desc_nn(a, b, both_dirs=True)
    torch.cuda.synchronize()
    row_idx, row_d2, col_idx, col_d2 = out[0], out[1], out[2], out[3]
    assert int(row_idx.min()) >= 0 and int(row_idx.max()) < n and int(col_idx.min()) >= 0
    # returned distance == fp64 distance of the returned index
    diff = a.double() - b[row_idx.long()].double()
    np.testing.assert_allclose(row_d2.double().cpu().numpy(), (diff * diff).sum(1).cpu().numpy(), rtol=1e-5, atol=1e-6)
    # a random sample of rows against the exact oracle (fp64 arg-min, first index)
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(n, 1500, replace=False))
    oi, od, od_second = odesc.desc_nn(a[rows].cpu().numpy(), b.cpu().numpy(), return_second=True)
    got = row_idx.cpu().numpy()[rows]
    tie = (od_second - od) <= 1e-6
    assert ((got == oi) | tie).all(), int(((got != oi) & ~tie).sum())
    # the other direction, same two checks
    diff = b.double() - a[col_idx.long()].double()
    np.testing.assert_allclose(col_d2.double().cpu().numpy(), (diff * diff).sum(1).cpu().numpy(), rtol=1e-5, atol=1e-6)
    cols = np.sort(rng.choice(n, 500, replace=False))
    oi, od, od_second = odesc.desc_nn(b[cols].cpu().numpy(), a.cpu().numpy(), return_second=True)
    got = col_idx.cpu().numpy()[cols]
    tie = (od_second - od) <= 1e-6
    assert ((got == oi) | tie).all(), int(((got != oi) & ~tie).sum())


def test_piecewise_icp_c1_tile_vs_oracle(cuda):
    """C1 at its full size: one 1 M-point tile pair (SURVEY 8(d) scene: independent epochs, known block motion) through
    f4l_piecewise_icp against oracle/piecewise.py -- same cells, same stable set, identical rows."""
    from fusion4landslide_b200 import ops, synth
    from oracle import piecewise as opw
    d = synth.make_scene(1_000_000, seed=21, device=cuda)
    src64, tgt64 = d["src"].double().contiguous(), d["tgt"].double().contiguous()
    dvfs, mag, counts, thr = ops.piecewise_icp(src64, tgt64, 5.0, 10)
    c = counts.tolist()
    o = opw.piecewise_icp(src64.cpu().numpy(), tgt64.cpu().numpy(), 5.0, 10)
    assert c[0] == o["dvfs"].shape[0] and c[0] > 100_000
    assert abs(thr.item() - o["thr"]) <= 1e-12 * max(1.0, abs(o["thr"]))
    np.testing.assert_array_equal(dvfs[:c[0]].cpu().numpy(), o["dvfs"])
    np.testing.assert_allclose(mag[:c[0]].cpu().numpy(), o["dvfms"][:, 3], rtol=4e-16, atol=0)     # fp64 norm: one ulp (fma)


def test_coarse2fine_c3_tile_sampled_vs_oracle(cuda, golden_dir):
    """C3 at its full tile size (625 k points per epoch, 3 superpoint levels) through the class entry point.  The
    oracle's O(N^2) descriptor search cannot run at this size, so the check is staged: descriptor labels of a random
    sample of voxels against the fp64 arg-min; voxel maps against the kd-tree except ties; and -- given the GPU's own
    correspondences -- the pair lists of every level and the fine matching of a sample of pairs per level against
    oracle/paths.c2f_tile; plus size-independent properties of the merged field."""
    import os
    from fusion4landslide_b200 import configs, nets, synth
    from fusion4landslide_b200.entry_c2f import Coarse2Fine
    from oracle import paths as opaths
    z = np.load(os.path.join(golden_dir, "nets_shipped.npz"))
    model = nets.ClusterFeatureNetWithAttention()
    model.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("agg/")})
    model = model.to(cuda).eval()
    w = {k[4:]: z[k] for k in z.files if k.startswith("agg/")}
    n = 625_000
    d = synth.make_scene(n, seed=31, device=cuda, desc_dim=64)
    ls = [d["labels_src"][k] for k in (1, 2, 3)]
    lt = [d["labels_tgt"][k] for k in (1, 2, 3)]
    tt = dict(src_pts=d["src"], tgt_pts=d["tgt"], partition_src=ls, partition_tgt=lt, feat_raw_src=d["src_feat"],
              feat_raw_tgt=d["tgt_feat"])
    c = Coarse2Fine(configs.fusion_config(tt, levels=[1, 2, 3], feat_aggregate_model=model))
    c.implement_c2f_matching()
    torch.cuda.synchronize()
    di, do = c.data_interim, c.data_output
    # descriptor labels: a sample of rows against the exact fp64 arg-min over ALL voxels of the other epoch
    fs, ft = di.tile_pts_sub_feat_src.cpu().numpy(), di.tile_pts_sub_feat_tgt.cpu().numpy()
    rng = np.random.default_rng(0)
    pick = rng.choice(fs.shape[0], 512, replace=False)
    oi, od, od2 = odesc.desc_nn(fs[pick], ft, return_second=True)
    lab = di.labels_from_3d.cpu().numpy()[pick]
    assert ((lab == oi) | ((od2 - od) <= 1e-6)).all()
    # the rest of the path on the GPU's correspondences: pair lists and sampled fine matching per level
    v2p = {k: di["idx_voxel2pts_" + k].cpu().numpy() for k in ("src", "tgt")}
    o = opaths.c2f_tile(d["src"].cpu().numpy(), d["tgt"].cpu().numpy(), [x.cpu().numpy() for x in ls],
                        [x.cpu().numpy() for x in lt], d["src_feat"].cpu().numpy(), d["tgt_feat"].cpu().numpy(), w,
                        voxel_size=float(c.method.voxel_size), median_max_resolution=float(np.float32(c.para.median_max_resolution)),
                        max_pairs_per_level=40, corr3d_given=di.corres_3d_voxel_from_3d_idx.cpu().numpy(), v2p_given=v2p)
    for k, raw in (("src", d["src"].cpu().numpy()), ("tgt", d["tgt"].cpu().numpy())):
        same = v2p[k] == o["idx_voxel2pts_kdtree_" + k]
        assert same.mean() > 0.97
        bad = np.nonzero(~same)[0][:2000]
        assert oknn.tie_rows(o[k + "_pts_sub"][bad], raw, 1).all()
    total_checked = 0
    for lv in range(3):
        ol = o["levels"][lv]
        k = len(ol["m"])
        _, spt_s = opaths.patch_lists(ls[lv].cpu().numpy(), 10)
        assert [int(x[0]) for x in do.spt_corres_src_multiple[lv][:k]] == [int(spt_s[a][0]) for a in ol["m"]]
        fr, of = c.fine_results_multiple[lv], ol["fine"]
        np.testing.assert_array_equal(fr.K[:k].cpu().numpy(), of["K"])
        np.testing.assert_array_equal(fr.status[:k].cpu().numpy(), of["status"])
        it = fr.iters[:k].cpu().numpy()
        ok = (of["status"] == 0) & (it == of["iters"])
        assert ((of["status"] == 0) & (it != of["iters"])).sum() <= 1
        Tg = fr.T[:k].cpu().numpy().astype(np.float64)
        src_np = d["src"].cpu().numpy().astype(np.float64)
        for q in np.nonzero(ok)[0]:
            if round(of["fitness"][q] * of["K"][q]) < 3:
                continue
            P = src_np[spt_s[ol["m"][q]]]
            To = of["T"][q].astype(np.float64)
            assert np.abs((P @ Tg[q][:3, :3].T + Tg[q][:3, 3]) - (P @ To[:3, :3].T + To[:3, 3])).max() < 1e-5
            total_checked += 1
    assert total_checked > 30
    merged = do.corres_3d_refine_apply_icp
    assert merged.shape[0] > 0.3 * n
    uniq = torch.unique(merged[:, :3], dim=0).shape[0]
    assert uniq == merged.shape[0]                       # only_3d: every source point at most once after the level merge
    n1 = do.corres_3d_refine_apply_icp_multiple[0].shape[0]
    assert torch.equal(merged[:n1], do.corres_3d_refine_apply_icp_multiple[0])          # level 1 first and complete
