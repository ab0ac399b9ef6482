"""GPU parity of the Python host mirror (same names / signatures as the reference's functions)
against the reference's golden outputs and the fp64 oracle.  These tests read like calls into the
reference: src.functions, scripts.weighted_svd, utils.o3d_tools, src.coarse_to_fine_matching(_base),
src.f2s3, src.piecewise_icp.  Tolerance: 1e-5 m on transformed points (north_star), indices exact."""
import os
import types

import numpy as np
import pytest
import torch
from scipy.spatial import cKDTree

from oracle import desc_nn as odesc
from oracle import icp as oicp
from oracle import knn as oknn
from oracle import piecewise as opw
from oracle import rigid
from oracle.make_golden import _patch
from tests.test_oracle_golden import _act_err, _cases

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_kabsch_transformation_estimation_signature_and_golden(cuda, golden_dir):
    from fusion4landslide_b200.functions import kabsch_transformation_estimation, transformation_residuals
    z = np.load(os.path.join(golden_dir, "rigid_kabsch.npz"))
    for k in _cases(z):
        s, t = z[k + "_src"], z[k + "_tgt"]
        w = None if z[k + "_w"].size == 0 else torch.from_numpy(z[k + "_w"]).to(cuda)[None]
        x1, x2 = torch.from_numpy(s).to(cuda)[None], torch.from_numpy(t).to(cuda)[None]
        R, tr, res, flag = kabsch_transformation_estimation(x1, x2, w)
        assert R.shape == (1, 3, 3) and tr.shape == (1, 3, 1) and res.shape == (1, s.shape[0]) and flag is False
        Ro, to, reso, _ = rigid.kabsch(s, t, None if w is None else z[k + "_w"])
        assert _act_err(R[0].cpu().numpy(), tr[0, :, 0].cpu().numpy(), Ro, to, s) < TOL
        np.testing.assert_allclose(res[0].cpu().numpy(), reso, atol=TOL)
        np.testing.assert_allclose(transformation_residuals(x1, x2, R, tr)[0].cpu().numpy(), reso, atol=2e-5)
    # batch of 3 problems at once, numpy inputs accepted
    rng = np.random.default_rng(0)
    ps = [_patch(rng, 50, rng.uniform(0, 50, 3)) for _ in range(3)]
    x1 = np.stack([p[0] for p in ps]).astype(np.float32)
    x2 = np.stack([p[1] for p in ps]).astype(np.float32)
    R, tr, res, _ = kabsch_transformation_estimation(x1, x2)
    for b in range(3):
        Ro, to, _, _ = rigid.kabsch(x1[b], x2[b], None)
        assert _act_err(R[b].cpu().numpy(), tr[b, :, 0].cpu().numpy(), Ro, to, x1[b]) < TOL


def test_weighted_procrustes_and_refine(cuda, golden_dir):
    from fusion4landslide_b200.weighted_svd import refine_local_rigid_correspondences, weighted_procrustes, weighted_svd
    z = np.load(os.path.join(golden_dir, "rigid_procrustes.npz"))
    for k in _cases(z)[:12]:
        s, t = z[k + "_src"], z[k + "_tgt"]
        w = None if z[k + "_w"].size == 0 else torch.from_numpy(z[k + "_w"]).to(cuda)
        S, Tg = torch.from_numpy(s).to(cuda), torch.from_numpy(t).to(cuda)
        T = weighted_procrustes(S, Tg, w, eps=1e-6)
        assert T.shape == (4, 4) and T.device.type == "cpu"            # quirk q1
        Ro, to = rigid.weighted_procrustes(s, t, None if w is None else z[k + "_w"], 0.0, 1e-6)
        assert _act_err(T[:3, :3].numpy(), T[:3, 3].numpy(), Ro, to, s) < TOL
        R, tv = weighted_procrustes(S[None], Tg[None], None if w is None else w[None], eps=1e-6, return_transform=False)
        assert R.shape == (1, 3, 3) and tv.shape == (1, 3) and R.is_cuda
        T2 = weighted_svd(S, Tg, weights=w)
        assert T2.is_cuda and _act_err(T2[:3, :3].cpu().numpy(), T2[:3, 3].cpu().numpy(), Ro, to, s) < TOL
    rng = np.random.default_rng(1)
    s, t = _patch(rng, 300, np.array([10.0, 20.0, 5.0]))
    t[::17] += 3.0                                                       # residuals > 1 m get pruned
    corr = np.hstack([s, t]).astype(np.float32)
    kept, T = refine_local_rigid_correspondences(torch.from_numpy(corr).to(cuda))
    ko, To = rigid.refine_local_rigid_correspondences(corr)
    assert kept.shape[0] == ko.shape[0] and T.is_cuda and T.dtype == torch.float32
    assert _act_err(T[:3, :3].cpu().numpy(), T[:3, 3].cpu().numpy(), To[:3, :3], To[:3, 3], s) < TOL


def test_rgb_guided_refine(cuda):
    """src/rgb_guided.py:99-125: res < 2.5 * median(res) mask and the 70 % quality flag."""
    from fusion4landslide_b200 import rgb_guided
    rng = np.random.default_rng(8)
    cases = []
    for n, frac in ((200, 0.05), (64, 0.45), (11, 0.0)):
        s, t = _patch(rng, n, rng.uniform(0, 60, 3), outliers=frac)
        cases.append(np.hstack([s, t]).astype(np.float32))
    ptr = np.zeros(len(cases) + 1, np.int32)
    ptr[1:] = np.cumsum([c.shape[0] for c in cases])
    allc = np.concatenate(cases)
    R, t, mask, mask2, res, med = rgb_guided.refine_local_rigid_correspondences_batched(
        torch.from_numpy(allc).to(cuda), torch.from_numpy(ptr).to(cuda))
    for i, c in enumerate(cases):
        Ro, to = rigid.weighted_procrustes(c[:, :3], c[:, 3:6], None, 0.0, 1e-6)
        ro = np.linalg.norm(c[:, :3].astype(np.float64) @ Ro.T + to - c[:, 3:6], axis=1)
        mo = ro < 2.5 * rigid.lower_median(ro)
        got = mask[ptr[i]:ptr[i + 1]].cpu().numpy()
        borderline = np.abs(ro - 2.5 * rigid.lower_median(ro)) < 1e-5
        assert ((got == mo) | borderline).all()
        assert bool(mask2[i]) == bool(mo.mean() >= 0.70) or abs(mo.mean() - 0.70) < 0.02
        assert _act_err(R[i].cpu().numpy(), t[i].cpu().numpy(), Ro, to, c[:, :3]) < TOL
        kept, T, m1, m2 = rgb_guided.refine_local_rigid_correspondences(torch.from_numpy(c).to(cuda))
        assert kept.shape[0] == int(got.sum()) and T.is_cuda and bool(m2) == bool(mask2[i])


def test_icp_registration_dict(cuda):
    from fusion4landslide_b200.o3d_tools import icp_registration
    rng = np.random.default_rng(2)
    s, t = _patch(rng, 180, np.array([30.0, 40.0, 3.0]))
    T0 = rigid.procrustes_transform(s, t, None, 0.0, 1e-6)
    pcd = types.SimpleNamespace(points=s.astype(np.float32))              # Open3D-like object
    out = icp_registration(pcd, t.astype(np.float32), T0, threshold=0.1)
    o = oicp.icp_point_to_point(s.astype(np.float32), t.astype(np.float32), T0, 0.1)
    assert set(out) == {"fitness", "inlier_rmse", "correspondence_set", "est_transform", "src_corr_pts", "tgt_corr_pts"}
    assert abs(out["fitness"] - o["fitness"]) < 1e-12 and abs(out["inlier_rmse"] - o["inlier_rmse"]) < 1e-7
    np.testing.assert_array_equal(out["correspondence_set"], o["correspondence_set"])
    assert _act_err(out["est_transform"][:3, :3], out["est_transform"][:3, 3], o["transformation"][:3, :3],
                    o["transformation"][:3, 3], s) < TOL
    with pytest.raises(ValueError):
        icp_registration(pcd, t, T0, icp_type="bogus")


def test_refine_dvfs_merge_c2c_transform(cuda, golden_dir):
    from fusion4landslide_b200 import coarse_to_fine as c2f
    from fusion4landslide_b200.functions import compute_c2c, transform_point_cloud
    z = np.load(os.path.join(golden_dir, "knn_sklearn.npz"))
    a, b = z["a"], z["b"]
    np.testing.assert_allclose(compute_c2c(a, b), z["c2c"], atol=2e-6)
    rng = np.random.default_rng(3)
    s = a[:5000]
    moved = (s + rng.normal(0, 0.03, s.shape)).astype(np.float32)
    rows = c2f.refine_dvfs_with_threshold(torch.from_numpy(s).to(cuda), torch.from_numpy(moved).to(cuda),
                                          torch.from_numpy(b).to(cuda), 0.05)
    ro, keep, _ = oknn.refine_dvfs_with_threshold(s, moved, b, 0.05)
    assert rows.shape == ro.shape and 0 < rows.shape[0] < 5000
    np.testing.assert_array_equal(rows.cpu().numpy(), ro)
    assert c2f.refine_dvfs_with_threshold(torch.zeros((0, 3), device=cuda), torch.zeros((0, 3), device=cuda),
                                          torch.from_numpy(b).to(cuda)).shape == (0, 6)
    # level merge: level 1 repeats 40 % of level 0 (within 1e-4 m) and adds new points
    l0 = np.hstack([a[:3000], a[:3000] + 0.1]).astype(np.float32)
    dup = a[:1200] + np.float32(5e-5)
    l1 = np.hstack([np.vstack([dup, a[3000:4500]]), np.zeros((2700, 3), np.float32)]).astype(np.float32)
    l2 = np.hstack([np.vstack([a[4000:5000], a[6000:6500]]), np.ones((1500, 3), np.float32)]).astype(np.float32)
    merged = c2f.merge_correspondences_by_priority_with_distance_threshold([torch.from_numpy(x).to(cuda) for x in (l0, l1, l2)])
    mo, _ = odesc.merge_by_priority([l0, l1, l2])
    np.testing.assert_array_equal(merged.cpu().numpy(), mo)
    # transform_point_cloud: numpy in -> numpy out, tensor in -> tensor out
    R = rigid.procrustes_transform(*_patch(rng, 30, np.zeros(3)))[:3, :3]
    tv = np.array([[0.5], [-0.25], [2.0]])
    out = transform_point_cloud(a[:100].astype(np.float64), R, tv)
    assert isinstance(out, np.ndarray)
    np.testing.assert_allclose(out, rigid.transform_point_cloud(a[:100], R, tv), atol=TOL)
    out_t = transform_point_cloud(torch.from_numpy(a[:100]).to(cuda), torch.from_numpy(R).float().to(cuda),
                                  torch.from_numpy(tv).float().to(cuda))
    assert out_t.is_cuda and np.abs(out_t.cpu().numpy() - out).max() < TOL
    # A2 voxel maps against scipy cKDTree (base.py:1038-1057)
    sub = a[::3] + np.float32(0.01)
    v2p, p2v = c2f.voxel_subsampling_maps(torch.from_numpy(sub).to(cuda), torch.from_numpy(a).to(cuda))
    _, ref = cKDTree(a.astype(np.float64)).query(sub.astype(np.float64))
    np.testing.assert_array_equal(v2p.cpu().numpy(), ref)
    p2v_ref = -np.ones(a.shape[0], np.int64)
    p2v_ref[ref] = np.arange(sub.shape[0])
    np.testing.assert_array_equal(p2v.cpu().numpy(), p2v_ref)


def test_coarse_matching_host(cuda, golden_dir):
    from fusion4landslide_b200 import coarse_to_fine as c2f
    z = np.load(os.path.join(golden_dir, "desc_cdist.npz"))
    cs, ct, fs, ft = z["coarse_cs"], z["coarse_ct"], z["coarse_fs"], z["coarse_ft"]
    mm = float(z["coarse_max_mag"][0])
    for kind in ("nn_mutual", "only_max_mag"):
        m, j = c2f.coarse_matching_3d(cs, fs, ct, ft, mm, kind)
        mo, jo = odesc.coarse_matching_3d(cs, fs, ct, ft, mm, kind)
        np.testing.assert_array_equal(m.cpu().numpy(), mo)
        np.testing.assert_array_equal(j.cpu().numpy(), jo)
    # 2D vote (B4)
    rng = np.random.default_rng(4)
    n_s, n_t, P = 20000, 21000, 150
    lab_s = rng.integers(0, P, n_s)
    lab_t = rng.integers(0, P + 20, n_t)
    corr2d = -np.ones((n_s, 2), np.int64)
    corr2d[:, 0] = np.arange(n_s)
    has = rng.random(n_s) < 0.3
    has[lab_s == 5] = False                                  # a patch without any 2D match
    corr2d[has, 1] = rng.integers(0, n_t, has.sum())
    order = np.argsort(lab_s, kind="stable")
    ptr = np.zeros(P + 1, np.int32)
    ptr[1:] = np.cumsum(np.bincount(lab_s, minlength=P))
    spt_src = [order[ptr[p]:ptr[p + 1]] for p in range(P)]
    idx_spt_tgt = np.array([l for l in range(P + 20) if l % 7 != 3])     # some tgt patches were removed
    T = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x)).to(cuda, dt)
    m, j, tie = c2f.coarse_matching_2d(T(corr2d, torch.int64), T(order, torch.int32), T(ptr, torch.int32),
                                       T(lab_t, torch.int64), T(idx_spt_tgt, torch.int64))
    mo, jo, tieo = odesc.coarse_matching_2d_vote(corr2d, lab_t, spt_src, idx_spt_tgt)
    # the oracle drops a patch whose (smallest-label) winner was removed; so does the kernel
    np.testing.assert_array_equal(m.cpu().numpy(), mo)
    np.testing.assert_array_equal(j.cpu().numpy(), jo)
    np.testing.assert_array_equal(tie.cpu().numpy(), tieo)


def test_f2s3_host(cuda, golden_dir):
    from fusion4landslide_b200 import f2s3
    rng = np.random.default_rng(6)
    n, m, D = 3000, 3300, 64
    fs = rng.standard_normal((n, D)).astype(np.float32)
    ft = rng.standard_normal((m, D)).astype(np.float32)
    fs /= np.linalg.norm(fs, axis=1, keepdims=True)
    ft /= np.linalg.norm(ft, axis=1, keepdims=True)
    sx, tx = rng.uniform(0, 50, (n, 3)).astype(np.float32), rng.uniform(0, 50, (m, 3)).astype(np.float32)
    labels, corr = f2s3.correspondence_searching(sx, tx, fs, ft)
    co, lo = odesc.f2s3_correspondences(sx, tx, fs, ft)
    np.testing.assert_array_equal(labels.cpu().numpy(), lo)
    np.testing.assert_array_equal(corr.cpu().numpy(), co)
    # pruning tail against the reference's own outputs (FilteringNetwork.filter_input, shipped weights)
    z = np.load(os.path.join(golden_dir, "f2s3_filter.npz"))
    keys = sorted({k.rsplit("_", 1)[0] for k in z.files if k.endswith("_corr")})
    for coeff, tag in ((1.0, "c1"), (2.5, "c25")):
        ks = [k for k in keys if k.endswith(tag)]
        corr = np.concatenate([z[k + "_corr"] for k in ks]).astype(np.float32)
        sc = np.concatenate([z[k + "_scores"] for k in ks]).astype(np.float32)
        ptr = np.zeros(len(ks) + 1, np.int32)
        ptr[1:] = np.cumsum([z[k + "_corr"].shape[0] for k in ks])
        R, t, robust, res = f2s3.filter_input_tail(torch.from_numpy(corr).to(cuda), torch.from_numpy(sc).to(cuda),
                                                   torch.from_numpy(ptr).to(cuda), coeff)
        for i, k in enumerate(ks):
            c = z[k + "_corr"]
            assert bool(robust[i]) == bool(z[k + "_robust"][0]), k
            assert _act_err(R[i].cpu().numpy(), t[i].cpu().numpy(), z[k + "_R"], z[k + "_t"], c[:, :3]) < 5e-5, k
            o = rigid.filter_input_tail(c[:, :3], c[:, 3:6], z[k + "_scores"], coeff)
            assert _act_err(R[i].cpu().numpy(), t[i].cpu().numpy(), o["rot_est"], o["trans_est"], c[:, :3]) < TOL, k
    # full pruning stage on the last batch: numpy restatement of f2s3.py:340-441
    rows, mag, keep = f2s3.correspondence_pruning(torch.from_numpy(corr).to(cuda), torch.from_numpy(sc).to(cuda),
                                                  torch.from_numpy(ptr).to(cuda), data_dir="x/Rockfall_Simulator/y",
                                                  refine_results=True, max_disp_magnitude=5.0, filter_median_magnitude=True)
    keep_o = sc > 0.99999
    for i, k in enumerate(ks):
        if bool(z[k + "_robust"][0]):
            keep_o[ptr[i]:ptr[i + 1]] = True
    r = corr[keep_o]
    mg = np.linalg.norm(r[:, :3] - r[:, 3:6], axis=1)
    r, mg = r[mg <= 5.0], mg[mg <= 5.0]
    sel = mg < 30 * np.median(mg)
    np.testing.assert_array_equal(keep.cpu().numpy(), keep_o)
    np.testing.assert_array_equal(rows.cpu().numpy(), r[sel])
    np.testing.assert_allclose(mag.cpu().numpy(), mg[sel], rtol=1e-6)


def test_piecewise_icp_entry_point(cuda, tmp_path):
    from fusion4landslide_b200.piecewise_icp import Piecewise_ICP
    from tests.test_piecewise_gpu import _scene
    src, tgt = _scene(40000, 5)
    np.save(tmp_path / "source_tile_0_overlap.npy", src)
    np.savetxt(tmp_path / "target_tile_0_overlap.txt", tgt, fmt="%.17g")
    cfg = types.SimpleNamespace(src_tile_overlap_path=str(tmp_path / "source_tile_0_overlap.npy"),
                                tgt_tile_overlap_path=str(tmp_path / "target_tile_0_overlap.txt"), smax=5.0,
                                number_points_min=10, threshold=0.1, output_root=str(tmp_path), tile_id=0,
                                dataset="brienz_tls", logging=None)
    assert Piecewise_ICP(cfg) is None
    o = opw.piecewise_icp(src, tgt, 5.0, 10)
    dvfs = np.loadtxt(tmp_path / "results" / "piecewise_icp_dvfs_of_tile_0.txt")
    dvfms = np.loadtxt(tmp_path / "results" / "piecewise_icp_dvfms_of_tile_0.txt")
    vis = np.loadtxt(tmp_path / "results" / "piecewise_dvfms_visualize_of_tile_0.txt")
    np.testing.assert_allclose(dvfs, o["dvfs"], atol=1e-9)
    np.testing.assert_allclose(dvfms, o["dvfms"], atol=1e-9)
    assert vis[0, 3] == 0 and vis[1, 3] == 5 and np.allclose(vis[2:], o["dvfms"][2:], atol=1e-9)


def test_host_pipeline_sparse_once_equals_doubled(cuda):
    """HostPipeline with the sparse rows crossing PCIe once (+ host expansion) returns the same tensors as the
    path that moves both copies."""
    from fusion4landslide_b200 import pipeline, synth
    tiles = []
    for s in range(3):
        d = synth.make_tile(30_000 + 5000 * s, seed=40 + s, device=cuda, patch_pts=200)
        tiles.append(pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"]))
    hts = [pipeline.HostTile(t) for t in tiles]
    a, _, da = pipeline.HostPipeline(hts, None, cuda, n_streams=2, sparse_once=False).run()
    b, _, db = pipeline.HostPipeline(hts, None, cuda, n_streams=2, sparse_once=True, expand_threads=3).run()
    assert db < da
    for x, y in zip(a, b):
        assert x["sparse"].shape == y["sparse"].shape and x["sparse"].shape[0] > 0
        for k in ("dense", "sparse", "T", "status", "median_resolution"):
            assert torch.equal(x[k], y[k]), k


def test_host_pipeline_compact_corr_equals_int64_tables(cuda):
    """HostPipeline(compact_corr=True): column 1 of the int64 correspondence tables repacked to int32 on the host
    (f4l_host_pack_corr_targets) and read through f4l_fine_buffers.corr3d_tgt -- same results, fewer bytes uploaded."""
    from fusion4landslide_b200 import ops, pipeline, synth
    tiles = []
    for s in range(3):
        d = synth.make_tile(30_000 + 5000 * s, seed=50 + s, device=cuda, patch_pts=200)
        tiles.append(pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"]))
    hts = [pipeline.HostTile(t) for t in tiles]
    a, ua, _ = pipeline.HostPipeline(hts, None, cuda, n_streams=2).run()
    b, ub, _ = pipeline.HostPipeline(hts, None, cuda, n_streams=2, compact_corr=True, pack_threads=2).run()
    assert ub < ua
    for x, y in zip(a, b):
        assert x["dense"].shape[0] > 0
        for k in ("dense", "sparse", "T", "status", "median_resolution"):
            assert torch.equal(x[k], y[k]), k
    # the host helper: out-of-range targets become -1
    c = torch.tensor([[0, 5], [1, -1], [2, 2 ** 40], [3, 2 ** 31 - 1], [4, -7]], dtype=torch.int64)
    assert ops.host_pack_corr_targets(c).tolist() == [5, -1, -1, 2 ** 31 - 1, -1]


def test_map_corr_2d_to_3d_golden(cuda, golden_dir):
    """2D match -> 3D point lifting (base.py:387-472) against the reference's own output.  Index / mask differences
    are allowed only where a hop is a documented tie: two candidates within 1e-6 (relative, squared pixel distance)
    or a hop length within 1e-4 px of the threshold (pixel coordinates are compared in f32)."""
    import os
    from scipy.spatial import cKDTree
    from fusion4landslide_b200 import coarse_to_fine as c2f
    z = np.load(os.path.join(golden_dir, "map_corr_2d.npz"))
    c, sp, tp, thr = z["corres_2d"], z["src_pixel"], z["tgt_pixel"], float(z["thres"][0])
    for rev, tag in ((False, "fwd"), (True, "rev")):
        fn = c2f.map_corr_2d_to_3d_tgt2src if rev else c2f.map_corr_2d_to_3d
        idx, mask, rows = fn(c, torch.from_numpy(sp).to(cuda), torch.from_numpy(tp).to(cuda), thr)
        idx, mask, rows = idx.cpu().numpy(), mask.cpu().numpy(), rows.cpu().numpy()
        a, b = (tp, sp) if rev else (sp, tp)
        ca, cb = (c[:, 2:4], c[:, :2]) if rev else (c[:, :2], c[:, 2:4])
        d1, _ = cKDTree(ca).query(a, k=2)
        hop = z["rows_" + tag][:, :2] if rev else z["rows_" + tag][:, 2:4]
        d2, _ = cKDTree(b).query(hop, k=2)
        tie = (d1[:, 1] ** 2 - d1[:, 0] ** 2 <= 1e-6 * d1[:, 1] ** 2) | (d2[:, 1] ** 2 - d2[:, 0] ** 2 <= 1e-6 * d2[:, 1] ** 2)
        edge = (np.abs(d1[:, 0] - thr) < 1e-4) | (np.abs(d2[:, 0] - thr) < 1e-4)
        same_rows = (rows == z["rows_" + tag]).all(1)
        assert (same_rows | tie).all()
        assert ((idx == z["idx_" + tag]) | tie).all()
        assert ((mask == z["mask_" + tag]) | tie | edge).all()
        assert tie.mean() < 1e-3 and edge.mean() < 1e-3


def test_displacement_field_side_stream_equals_serial(cuda):
    """A1 on a side stream (median_ready_event in the C ABI: the library waits in-stream before the assign kernel)
    gives bit-identical results to the serial order, also with many tiles over several streams."""
    from fusion4landslide_b200 import pipeline, synth
    tiles = []
    for s in range(4):
        d = synth.make_tile(40_000, seed=60 + s, device=cuda, patch_pts=220)
        tiles.append(pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"]))
    base = [pipeline.displacement_field(t) for t in tiles]
    torch.cuda.synchronize()
    streams, sides = pipeline.make_streams(2, cuda), pipeline.make_streams(2, cuda)
    got = pipeline.displacement_field_tiles(tiles, streams=streams, side_streams=sides)
    torch.cuda.synchronize()
    for (r0, m0), (r1, m1) in zip(base, got):
        assert m0.item() == m1.item()
        d0, s0, _ = r0.rows()
        d1, s1, _ = r1.rows()
        assert torch.equal(d0, d1) and torch.equal(s0, s1) and torch.equal(r0.T64, r1.T64)
