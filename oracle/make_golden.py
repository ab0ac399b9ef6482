"""TEST INFRASTRUCTURE -- writes tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container (where /root/reference is mounted):
    python -m oracle.make_golden
The reference modules are imported from where they lie through oracle/ref_shim.py (stand-ins
for Open3D & co. that the functions used here never touch) and executed on CPU torch (fp32, as
the reference computes).  The fixtures travel to the GPU box; the reference does not.

What gets pinned (reference function -> golden file):
  scripts/weighted_svd.py weighted_procrustes                      -> rigid_procrustes.npz
  src/functions.py kabsch_transformation_estimation, transformation_residuals,
      transform_point_cloud                                        -> rigid_kabsch.npz
  src/models/outlier_classifier.py FilteringNetwork.filter_input with the shipped weights
      (weights/outlier_classifier_best.pt)                         -> f2s3_filter.npz
  src/functions.py compute_c2c (sklearn kd_tree) + the k=2 median-resolution call pattern of
      base.py:2727-2736                                            -> knn_sklearn.npz
  torch.cdist + min call pattern of base.py:2805-2815 / 2966-2986  -> desc_cdist.npz
  torch.cdist rigidity expression of base.py:3310-3317             -> rigidity_cdist.npz
  src/data_loader.py Preprocess_Dataset.extract_patch, with a cKDTree stand-in for the one Open3D
      class it uses (KDTreeFlann.search_radius_vector_3d)          -> dips_patches.npz
  src/coarse_to_fine_matching_base.py map_corr_2d_to_3d, map_corr_2d_to_3d_tgt2src
      (scipy cKDTree, present here)                                -> map_corr_2d.npz
  src/coarse_to_fine_matching_base.py Coarse2Fine_Base.coarse_matching_with_different_types, the method body
      itself on a stand-in `self` (2D vote, mutual 3D, pair order)  -> coarse_method.npz
  src/coarse_to_fine_matching.py merge_correspondences_by_priority_with_distance_threshold with an EXACT
      stand-in for the faiss HNSW index                            -> merge_levels.npz
  FilteringNetwork.compute_weights / ClusterFeatureNetWithAttention.aggregation with the shipped weights
      (per-patch loops of f2s3.py:340-347, base.py:2561-2656)      -> nets_shipped.npz
Open3D ICP / Octree, hnswlib and faiss' HNSW cannot be run anywhere here -> no golden, parity unpinned.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_shim  # noqa: E402


def _rand_rigid(rng, max_deg=3.0, max_t=0.5):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.radians(rng.uniform(0, max_deg))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    t = rng.normal(size=3)
    t *= rng.uniform(0, max_t) / np.linalg.norm(t)
    return R, t


def _patch(rng, n, centre, extent=2.0, noise=0.01, outliers=0.1, planar=0.3):
    p = rng.uniform(-extent, extent, size=(n, 3))
    p[:, 2] *= planar
    p += centre
    R, t = _rand_rigid(rng)
    c = p.mean(0)
    q = (p - c) @ R.T + c + t + noise * rng.normal(size=(n, 3))
    bad = rng.random(n) < outliers
    q[bad] += rng.normal(size=(int(bad.sum()), 3)) * 0.5
    return p.astype(np.float32), q.astype(np.float32)


def make_rigid():
    rng = np.random.default_rng(0)
    ws = ref_shim.ref_weighted_svd()
    fn = ref_shim.ref_functions()
    cases_p, cases_k = {}, {}
    sizes = [3, 4, 10, 17, 64, 256, 1000, 4096]
    for ci, n in enumerate(sizes):
        centre = rng.uniform(-40, 40, size=3)
        s, t = _patch(rng, n, centre)
        for wi, wkind in enumerate(["none", "rand", "binary"]):
            if wkind == "none":
                w = None
            elif wkind == "rand":
                w = rng.random(n).astype(np.float32)
            else:
                w = (rng.random(n) < 0.6).astype(np.float32)
                w[:3] = 1
            key = "c%d_w%d" % (ci, wi)
            st, tt = torch.from_numpy(s), torch.from_numpy(t)
            wt = None if w is None else torch.from_numpy(w)
            # D1 / D2 call pattern (weighted_svd.py:134-142: eps=1e-6, return_transform=False)
            R, tr = ws.weighted_procrustes(st, tt, weights=wt, weight_thresh=0.0, eps=1e-6,
                                           return_transform=False, return_rmse=True)
            T = ws.weighted_procrustes(st, tt, weights=wt, eps=1e-7)          # defaults -> 4x4
            cases_p[key + "_src"] = s
            cases_p[key + "_tgt"] = t
            cases_p[key + "_w"] = np.zeros(0, np.float32) if w is None else w
            cases_p[key + "_R"] = R.numpy()
            cases_p[key + "_t"] = tr.numpy()
            cases_p[key + "_T_default"] = T.numpy()
            # D3 call pattern (outlier_classifier.py:73: defaults, b=1)
            Rk, tk, res, flag = fn.kabsch_transformation_estimation(st[None], tt[None], None if wt is None else wt[None])
            cases_k[key + "_src"] = s
            cases_k[key + "_tgt"] = t
            cases_k[key + "_w"] = np.zeros(0, np.float32) if w is None else w
            cases_k[key + "_R"] = Rk[0].numpy()
            cases_k[key + "_t"] = tk[0, :, 0].numpy()
            cases_k[key + "_res"] = res[0].numpy()
            cases_k[key + "_flag"] = np.array([flag])
            x1t = fn.transform_point_cloud(st, Rk[0], tk[0])                # D5
            cases_k[key + "_x1t"] = x1t.numpy()
    # reflection case: tgt is a mirrored copy -> det fix must trigger
    s, _ = _patch(rng, 50, np.zeros(3), planar=1.0)
    t = s.copy()
    t[:, 2] *= -1
    R, tr = ws.weighted_procrustes(torch.from_numpy(s), torch.from_numpy(t), eps=1e-6, return_transform=False)
    cases_p.update(refl_src=s, refl_tgt=t, refl_w=np.zeros(0, np.float32), refl_R=R.numpy(), refl_t=tr.numpy(),
                   refl_T_default=np.eye(4, dtype=np.float32))
    np.savez_compressed(os.path.join(GOLD, "rigid_procrustes.npz"), **cases_p)
    np.savez_compressed(os.path.join(GOLD, "rigid_kabsch.npz"), **cases_k)


def make_f2s3_filter():
    """The F2S3 pruning loop body (f2s3.py:340-347) on synthetic supervoxels through the real
    FilteringNetwork with the shipped weights."""
    rng = np.random.default_rng(1)
    oc = ref_shim.ref_outlier_classifier()
    net = oc.FilteringNetwork()
    sd = torch.load(os.path.join(ref_shim.REFERENCE_ROOT, "weights", "outlier_classifier_best.pt"),
                    map_location="cpu", weights_only=False)
    sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd
    missing = net.load_state_dict(sd, strict=False)
    net.eval()

    class Cfg:
        data_dir = "synthetic"
    out = {"load_report": np.array([str(missing)])}
    with torch.no_grad():
        for ci, n in enumerate([12, 40, 150, 400, 1200]):
            centre = rng.uniform(-30, 30, size=3)
            s, t = _patch(rng, n, centre, outliers=0.2 if ci % 2 else 0.05)
            svl = torch.from_numpy(np.concatenate([s, t], 1).astype(np.float64))
            scaled = torch.divide(svl, torch.max(torch.abs(svl)))
            for coeff_name, cfg_dir in (("c1", "synthetic"), ("c25", "Rockfall_Simulator")):
                Cfg.data_dir = cfg_dir
                o = net.filter_input(scaled.unsqueeze(0).unsqueeze(0).float(), svl.unsqueeze(0).float(), Cfg)
                key = "s%d_%s" % (ci, coeff_name)
                out[key + "_corr"] = svl.float().numpy()
                out[key + "_scores"] = o["scores"].reshape(-1).numpy()
                out[key + "_R"] = o["rot_est"].numpy()
                out[key + "_t"] = o["trans_est"].reshape(3).numpy()
                out[key + "_robust"] = np.array([bool(o["robust_estimate"])])
    np.savez_compressed(os.path.join(GOLD, "f2s3_filter.npz"), **out)


def make_knn():
    from sklearn.neighbors import NearestNeighbors
    fn = ref_shim.ref_functions()
    rng = np.random.default_rng(2)
    n = 20000
    xy = rng.uniform(0, 14, size=(n, 2))
    z = 1.5 * np.sin(xy[:, 0] * 0.7) + 0.8 * np.cos(xy[:, 1] * 1.3)
    a = np.c_[xy, z].astype(np.float32)
    xy2 = rng.uniform(0, 14, size=(n + 777, 2))
    z2 = 1.5 * np.sin(xy2[:, 0] * 0.7) + 0.8 * np.cos(xy2[:, 1] * 1.3) + 0.02
    b = np.c_[xy2, z2].astype(np.float32)
    # base.py:2727-2736 call pattern
    neigh = NearestNeighbors(n_neighbors=2, algorithm="kd_tree")
    neigh.fit(a)
    d_a, i_a = neigh.kneighbors(a, return_distance=True)
    neigh.fit(b)
    d_b, i_b = neigh.kneighbors(b, return_distance=True)
    res = max(np.median(d_a[:, -1]), np.median(d_b[:, -1]))
    c2c = fn.compute_c2c(a.astype(np.float64), b.astype(np.float64))       # functions.py:127-144
    neigh1 = NearestNeighbors(n_neighbors=1, algorithm="kd_tree").fit(b.astype(np.float64))
    _, c2c_idx = neigh1.kneighbors(a.astype(np.float64))
    np.savez_compressed(os.path.join(GOLD, "knn_sklearn.npz"), a=a, b=b, self_d_a=d_a, self_i_a=i_a,
                        self_d_b=d_b, self_i_b=i_b, median_resolution=np.array([res]),
                        c2c=c2c, c2c_idx=c2c_idx)


def make_desc():
    rng = np.random.default_rng(3)
    out = {}
    for D in (32, 64):
        n, m = 3000, 3500
        fs = rng.normal(size=(n, D)).astype(np.float32)
        fs /= np.linalg.norm(fs, axis=1, keepdims=True)
        ft = np.vstack([fs + 0.15 * rng.normal(size=(n, D)).astype(np.float32),
                        rng.normal(size=(m - n, D)).astype(np.float32)])
        ft /= np.linalg.norm(ft, axis=1, keepdims=True)
        ft = ft[rng.permutation(m)].astype(np.float32)
        a, b = torch.from_numpy(fs), torch.from_numpy(ft)
        labels = torch.empty(n, dtype=torch.long)
        dists = torch.empty(n)
        for i in range(0, n, 1024):                                        # base.py:2805-2815
            d = torch.cdist(a[i:i + 1024], b)
            md, mi = d.min(dim=1)
            labels[i:i + 1024] = mi
            dists[i:i + 1024] = md
        out["D%d_a" % D] = fs
        out["D%d_b" % D] = ft
        out["D%d_labels" % D] = labels.numpy()
        out["D%d_dist" % D] = dists.numpy()
    # coarse matching expression, base.py:2966-2986
    S, T, D = 400, 380, 64
    cs = rng.uniform(0, 60, size=(S, 3)).astype(np.float32)
    perm = rng.permutation(S)[:T]
    ct = (cs[perm] + rng.normal(size=(T, 3)) * 0.3).astype(np.float32)
    fs = rng.normal(size=(S, D)).astype(np.float32)
    fs /= np.linalg.norm(fs, axis=1, keepdims=True)
    ft = fs[perm] + 0.3 * rng.normal(size=(T, D)).astype(np.float32)
    ft /= np.linalg.norm(ft, axis=1, keepdims=True)
    ft = ft.astype(np.float32)
    max_mag = 5.0
    dc = torch.cdist(torch.from_numpy(cs), torch.from_numpy(ct))
    df = torch.cdist(torch.from_numpy(fs), torch.from_numpy(ft))
    df[dc > max_mag] = torch.inf
    d_t = torch.min(df, dim=1)
    d_s = torch.min(df, dim=0)
    mutual = torch.zeros(S, dtype=torch.bool)
    for m, i in enumerate(d_t[1]):
        if d_s[1][i] == m:
            mutual[m] = True
    in_mag = d_t[0] < torch.inf
    out.update(coarse_cs=cs, coarse_ct=ct, coarse_fs=fs, coarse_ft=ft, coarse_max_mag=np.array([max_mag]),
               coarse_j=d_t[1].numpy(), coarse_mutual=mutual.numpy(), coarse_in_mag=in_mag.numpy())
    np.savez_compressed(os.path.join(GOLD, "desc_cdist.npz"), **out)


def make_rigidity():
    rng = np.random.default_rng(4)
    out = {}
    for ci, n in enumerate([10, 25, 26, 100, 500]):
        s, t = _patch(rng, n, rng.uniform(-40, 40, size=3), outliers=0.3 if ci % 2 else 0.05)
        A, B = torch.from_numpy(s), torch.from_numpy(t)
        dS = torch.cdist(A, A, p=2)                                        # base.py:3310-3317
        dT = torch.cdist(B, B, p=2)
        diff = torch.abs(dS - dT)
        num_ele = len(diff) * (len(diff) - 1) / 2
        dist_mean = torch.sum(torch.triu(diff, diagonal=1)) / num_ele
        ratio = (torch.sum(diff <= 0.5) - diff.shape[0]) / (num_ele * 2)
        out["r%d_src" % ci] = s
        out["r%d_tgt" % ci] = t
        out["r%d_ratio" % ci] = np.array([float(ratio)])
        out["r%d_mean" % ci] = np.array([float(dist_mean)])
    np.savez_compressed(os.path.join(GOLD, "rigidity_cdist.npz"), **out)


def make_dips():
    """src/data_loader.py run unmodified: the module-level `o3d` name is pointed at a stand-in whose
    KDTreeFlann answers radius queries through oracle.dips.radius_search (nanoflann semantics)."""
    import importlib
    import types
    from scipy.spatial import cKDTree
    from oracle import dips as odips
    ref_shim.load()
    dl = importlib.import_module("src.data_loader")

    class _Pcd:
        def __init__(self, pts):
            self.points = pts

    class _Tree:
        def __init__(self, pcd):
            self.pts = np.asarray(pcd.points)
            self.tree = cKDTree(self.pts)
            self.sizes = []

        def search_radius_vector_3d(self, pt, radius):
            idx, d2 = odips.radius_search(self.tree, self.pts, np.asarray(pt), radius)
            self.sizes.append(idx.size)
            return idx.size, idx.tolist(), d2.tolist()

    dl.o3d = types.SimpleNamespace(geometry=types.SimpleNamespace(KDTreeFlann=_Tree))
    rng = np.random.default_rng(21)
    n = 7000
    xy = rng.uniform(0, 8.4, (n, 2))                        # ~100 pts / m^2 like a 0.1 m voxel grid
    z = 0.8 * np.sin(xy[:, 0] * 0.9) + 0.5 * np.cos(xy[:, 1] * 1.3) + 0.01 * rng.normal(size=n)
    ref = np.column_stack([xy, z])
    lonely = np.array([[30.0, 30.0, 1.0]]) + 0.05 * rng.normal(size=(7, 3))      # <= 10 neighbours: no frame
    wall = np.column_stack([rng.uniform(40, 41.2, 300), np.full(300, 40.0) + 0.005 * rng.normal(size=300),
                            rng.uniform(0, 1.2, 300)])                          # small vertical patch (< 256 pts)
    mid = np.array([[60.0, 10.0, 3.0]]) + 0.3 * rng.normal(size=(60, 3))         # 10 < n < 256: frame + zero padding
    ref = np.vstack([ref, lonely, wall, mid])                # tile-local coordinates
    radius = float(np.sqrt(3) * 10 * 0.1)                    # f2s3.py:106 with a 0.1 m median resolution: ~940 neighbours
    pick = np.concatenate([rng.choice(n, 36, replace=False), n + np.arange(7), n + 7 + rng.choice(300, 7, replace=False),
                           n + 307 + rng.choice(60, 6, replace=False)])
    data = ref[pick]
    ds = dl.Preprocess_Dataset(_Pcd(data), _Pcd(ref), points_per_batch=len(pick), feature_radius=radius)
    np.random.seed(1234)
    out = ds[0].numpy()                                      # (n_pick, 3, 256) f32
    sizes = list(ds.pcd_tree.sizes)
    np.random.seed(1234)
    inds = np.stack([np.random.choice(max(s_, 256), 256, replace=False) for s_ in sizes]).astype(np.int32)
    mine, cnt, lrf = odips.patches(data, ref, radius, inds)
    assert (cnt == np.asarray(sizes)).all()
    err = np.abs(mine - out).max()
    print("dips: %d queries, neighbours %d..%d, oracle vs reference max |diff| = %.2e" % (len(pick), min(sizes), max(sizes), err))
    assert err < 2e-6
    np.savez_compressed(os.path.join(GOLD, "dips_patches.npz"), ref=ref, pick=pick.astype(np.int32), radius=np.array([radius]),
                        inds=inds, patches=out, count=np.asarray(sizes, np.int32), lrf=lrf)


def make_lifting():
    """src/coarse_to_fine_matching_base.py map_corr_2d_to_3d / map_corr_2d_to_3d_tgt2src, unmodified."""
    base = ref_shim.ref_base()
    from oracle import lifting as olift
    rng = np.random.default_rng(33)
    W, H = 640.0, 800.0                                   # a crop: keeps the fixture small at realistic densities
    n_match, n_src, n_tgt = 4000, 3000, 8000
    m_src = rng.uniform(0, [W, H], (n_match, 2))
    m_tgt = m_src + rng.normal(0, 6.0, (n_match, 2)) + np.array([12.0, -7.0])
    corres_2d = np.round(np.hstack([m_src, m_tgt]), 3)                       # np.loadtxt of "%.3f" files
    src_pixel = torch.from_numpy(rng.uniform(0, [W, H], (n_src, 2)).astype(np.float32))
    tgt_pixel = torch.from_numpy((rng.uniform(0, [W, H], (n_tgt, 2))).astype(np.float32))
    thres = 5.0
    i_f, m_f, r_f = base.map_corr_2d_to_3d(corres_2d, src_pixel, tgt_pixel, thres)
    i_r, m_r, r_r = base.map_corr_2d_to_3d_tgt2src(corres_2d, src_pixel, tgt_pixel, thres)
    oi, om, orow = olift.map_corr_2d_to_3d(corres_2d, src_pixel.numpy(), tgt_pixel.numpy(), thres)
    assert np.array_equal(oi, i_f) and np.array_equal(om, m_f) and np.array_equal(orow, r_f)
    oi, om, orow = olift.map_corr_2d_to_3d(corres_2d, src_pixel.numpy(), tgt_pixel.numpy(), thres, reverse=True)
    assert np.array_equal(oi, i_r) and np.array_equal(om, m_r) and np.array_equal(orow, r_r)
    print("lifting: %d / %d source points and %d / %d target points lifted; oracle == reference" %
          (int(m_f.sum()), n_src, int(m_r.sum()), n_tgt))
    np.savez_compressed(os.path.join(GOLD, "map_corr_2d.npz"), corres_2d=corres_2d, src_pixel=src_pixel.numpy(),
                        tgt_pixel=tgt_pixel.numpy(), thres=np.array([thres]), idx_fwd=i_f.astype(np.int64), mask_fwd=m_f,
                        rows_fwd=r_f, idx_rev=i_r.astype(np.int64), mask_rev=m_r, rows_rev=r_r)


class _Log:
    def info(self, *a, **k):
        pass


def make_coarse():
    """B3 + B4 + the pair ordering: the reference's OWN method body `Coarse2Fine_Base.coarse_matching_with_different_types`
    (base.py:2925-3157) executed unmodified on a stand-in `self` (fusion mode: 2D-vote pairs followed by mutual 3D
    pairs; the learned superpoint features are handed in, so `_compute_spt_feat_and_coord_with_fused_feats` is a no-op).
    Source patches whose best vote count is shared by several target labels are recorded (`tie`): torch.argsort's order
    of equal counts is unspecified, the golden holds what torch-CPU returned."""
    base = ref_shim.ref_base()
    ED = sys.modules["easydict"].EasyDict
    rng = np.random.default_rng(11)
    n_s, n_t, P = 6000, 6400, 60
    # patch labels: contiguous blocks with gaps in the label space, a few patches below the size gate (removed)
    lab_s = np.sort(rng.integers(0, P, n_s)) * 3 + 5
    lab_t = np.sort(rng.integers(0, P, n_t)) * 3 + 5
    lab_t[lab_t == 5 + 3 * 7] = 5 + 3 * 8                       # one target label never occurs
    perm_s, perm_t = rng.permutation(n_s), rng.permutation(n_t)
    lab_s, lab_t = lab_s[perm_s], lab_t[perm_t]

    def lists(lab, min_pts):
        u, c = np.unique(lab, return_counts=True)
        keep = u[c > min_pts]
        return keep, [torch.from_numpy(np.nonzero(lab == k)[0]) for k in keep]

    idx_spt_src, spt2pts_src = lists(lab_s, 40)
    idx_spt_tgt, spt2pts_tgt = lists(lab_t, 40)
    # 2D-lifted matches: 30 % of the source points, most into the "same" patch id, some elsewhere, some ties by design
    corr2d = -np.ones((n_s, 2), np.int64)
    corr2d[:, 0] = np.arange(n_s)
    pts_of_t = {int(k): np.nonzero(lab_t == k)[0] for k in np.unique(lab_t)}
    keys_t = np.array(sorted(pts_of_t))
    for i in np.nonzero(rng.random(n_s) < 0.3)[0]:
        k = int(lab_s[i])
        if rng.random() < 0.35 or k not in pts_of_t:
            k = int(keys_t[rng.integers(0, keys_t.size)])
        corr2d[i, 1] = rng.choice(pts_of_t[k])
    # force exact ties in three source patches: two matches each into two target patches
    for m in (3, 11, 20):
        pts = spt2pts_src[m].numpy()
        corr2d[pts, 1] = -1
        ka, kb = int(keys_t[2 * m]), int(keys_t[2 * m + 1])
        corr2d[pts[0], 1], corr2d[pts[1], 1] = pts_of_t[ka][0], pts_of_t[ka][1]
        corr2d[pts[2], 1], corr2d[pts[3], 1] = pts_of_t[kb][0], pts_of_t[kb][1]
    # superpoint coordinates / features for the 3D branch
    S, T, D = len(spt2pts_src), len(spt2pts_tgt), 64
    cs = rng.uniform(0, 60, size=(S, 3)).astype(np.float32)
    ct = (cs[rng.integers(0, S, T)] + rng.normal(size=(T, 3)) * 0.4).astype(np.float32)
    fs = rng.normal(size=(S, D)).astype(np.float32)
    fs /= np.linalg.norm(fs, axis=1, keepdims=True)
    ft = (fs[rng.integers(0, S, T)] + 0.4 * rng.normal(size=(T, D))).astype(np.float32)
    ft /= np.linalg.norm(ft, axis=1, keepdims=True)
    out = dict(lab_s=lab_s, lab_t=lab_t, corr2d=corr2d, cs=cs, ct=ct, fs=fs, ft=ft, min_pts=np.array([40]),
               max_mag=np.array([5.0]))

    for mode in ("fusion", "only_2d", "only_3d"):
        class Fake:
            pass
        me = Fake()
        me.verbose, me.logging, me.device = False, _Log(), "cpu"
        me.method = ED(coarse_matching_fusion=mode == "fusion", coarse_matching_only_3d=mode == "only_3d",
                       coarse_matching_only_2d=mode == "only_2d", fine_matching_only_2d=False,
                       feat_aggregate_type="learning_based", use_img_patch_enhanced_3d_aggregation=False,
                       use_img_pixel_enhanced_3d_aggregation=False, use_normal_3d_aggregation=True,
                       coarse_refinement_3d_type="nn_mutual")
        me.para = ED(max_magnitude=5.0)
        me.visualize = ED(visualize_patch=False)
        me.data_interim = ED()
        di = me.data_interim
        di.idx_spt2pts_src, di.idx_spt2pts_tgt = list(spt2pts_src), list(spt2pts_tgt)
        di.idx_spt_src, di.idx_spt_tgt = torch.from_numpy(idx_spt_src), torch.from_numpy(idx_spt_tgt)
        di.idx_pts2spt_src, di.idx_pts2spt_tgt = torch.from_numpy(lab_s), torch.from_numpy(lab_t)
        di.corres_3d_from_2d_idx = torch.from_numpy(corr2d)
        di.spt_coord_src, di.spt_coord_tgt = torch.from_numpy(cs), torch.from_numpy(ct)
        di.spt_feat_src, di.spt_feat_tgt = torch.from_numpy(fs), torch.from_numpy(ft)
        me.data_output = ED()
        me._compute_spt_feat_and_coord_with_fused_feats = lambda: None
        base.Coarse2Fine_Base.coarse_matching_with_different_types(me)
        # identify every returned list by its first point index -> patch position
        first_s = {int(x[0]): k for k, x in enumerate(spt2pts_src)}
        first_t = {int(x[0]): k for k, x in enumerate(spt2pts_tgt)}
        m = np.array([first_s[int(x[0])] for x in me.data_output.spt_corres_src], np.int64)
        j = np.array([first_t[int(x[0])] for x in me.data_output.spt_corres_tgt], np.int64)
        out[mode + "_m"], out[mode + "_j"] = m, j
        out[mode + "_spt_length"] = np.array(me.spt_length, np.int64)
    # which source patches have a tied top count (independent of the reference: plain counting)
    tie = np.zeros(S, bool)
    for m, pts in enumerate(spt2pts_src):
        t = corr2d[pts.numpy(), 1]
        t = t[t >= 0]
        if t.size:
            _, c = np.unique(lab_t[t], return_counts=True)
            tie[m] = (c == c.max()).sum() > 1
    out["tie"] = tie
    np.savez_compressed(os.path.join(GOLD, "coarse_method.npz"), **out)


def make_merge():
    """M1: the reference's OWN `merge_correspondences_by_priority_with_distance_threshold`
    (src/coarse_to_fine_matching.py:40-118, default search_type='faiss') executed unmodified, with `faiss` replaced by an
    EXACT stand-in index (brute force, float32 squared L2 like faiss' fvec_L2sqr): the result is what the function
    returns when its HNSW index (approximate, absent here) answers every query correctly.  The function's 'cdist' and
    'kdtree' branches are broken in the reference (torch.cat over numpy arrays / inverted mask) and cannot be run."""
    ref_shim.load()
    import types

    class _Hnsw:
        efConstruction = 0
        efSearch = 0

    class IndexHNSWFlat:
        def __init__(self, dim, m):
            self.dim, self.hnsw, self.x = dim, _Hnsw(), np.zeros((0, dim), np.float32)

        def add(self, x):
            self.x = np.concatenate([self.x, np.asarray(x, np.float32).reshape(-1, self.dim)], 0)

        def search(self, q, k):
            assert k == 1
            q = np.asarray(q, np.float32)
            D = np.full((q.shape[0], 1), np.float32(np.inf), np.float32)
            I = np.full((q.shape[0], 1), -1, np.int64)
            for s in range(0, q.shape[0], 512):
                d = ((q[s:s + 512, None, :] - self.x[None, :, :]) ** 2).sum(-1, dtype=np.float32)
                if d.shape[1]:
                    I[s:s + 512, 0] = d.argmin(1)
                    D[s:s + 512, 0] = d.min(1)
            return D, I

    fake = types.ModuleType("faiss")
    fake.IndexHNSWFlat = IndexHNSWFlat
    sys.modules["faiss"] = fake
    import importlib
    mod = importlib.import_module("src.coarse_to_fine_matching")
    mod.faiss = fake
    rng = np.random.default_rng(12)
    src = rng.uniform(0, 40, size=(5000, 3)).astype(np.float32)
    out = {}
    levels = []
    for lv, frac in enumerate((0.5, 0.6, 0.7)):
        pick = np.nonzero(rng.random(src.shape[0]) < frac)[0]
        rows = np.concatenate([src[pick], src[pick] + rng.normal(size=(pick.size, 3)).astype(np.float32) * 0.1], 1)
        if lv == 2:                                   # near-duplicates on both sides of the 1e-3 threshold
            rows[:50, 0] += np.float32(5e-4)
            rows[50:100, 0] += np.float32(2e-3)
        levels.append(rows.astype(np.float32))
        out["level%d" % lv] = levels[-1]
    merged = mod.merge_correspondences_by_priority_with_distance_threshold([torch.from_numpy(x) for x in levels])
    out["merged"] = merged.numpy()
    levels_e = [levels[0][:0], levels[1], levels[2]]     # empty first level
    out["merged_empty0"] = mod.merge_correspondences_by_priority_with_distance_threshold(
        [torch.from_numpy(x) for x in levels_e]).numpy()
    np.savez_compressed(os.path.join(GOLD, "merge_levels.npz"), **out)


def make_nets():
    """8(f)-2: the reference's two small learned models with the SHIPPED weights, executed unmodified per patch:
      FilteringNetwork.compute_weights (src/models/outlier_classifier.py:52-63) per supervoxel, scaled as f2s3.py:343-346
      ClusterFeatureNetWithAttention.aggregation, mode 'test' (cluster_feature_net_self_attention.py:72-103)
    The weights travel with the golden (float32 arrays keyed by state_dict key) so the GPU test runs the batched
    kernels with the same parameters."""
    ref_shim.load()
    import importlib
    ED = sys.modules["easydict"].EasyDict
    rng = np.random.default_rng(13)
    out = {}
    oc = ref_shim.ref_outlier_classifier()
    net = oc.FilteringNetwork()
    sd = torch.load(os.path.join(ref_shim.REFERENCE_ROOT, "weights", "outlier_classifier_best.pt"), map_location="cpu",
                    weights_only=False)
    sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd
    net.load_state_dict(sd)
    net.eval()
    for k, v in net.state_dict().items():
        out["filter/" + k] = v.numpy()
    sizes = [11, 12, 37, 150, 33, 400, 64, 1200, 15]
    rows, scores = [], []
    with torch.no_grad():
        for ci, n in enumerate(sizes):
            s, t = _patch(rng, n, rng.uniform(-30, 30, size=3), outliers=0.25 if ci % 2 else 0.05)
            svl = torch.from_numpy(np.concatenate([s, t], 1).astype(np.float64))
            scaled = torch.divide(svl, torch.max(torch.abs(svl)))                       # f2s3.py:343
            w = net.compute_weights(scaled.unsqueeze(0).unsqueeze(0).float())            # :346
            rows.append(svl.numpy())
            scores.append(w.reshape(-1).numpy())
    out["filter_corr"] = np.concatenate(rows, 0)
    out["filter_ptr"] = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    out["filter_scores"] = np.concatenate(scores, 0)

    cf = importlib.import_module("src.feature_aggregation.cluster_feature_net_self_attention")
    model = cf.ClusterFeatureNetWithAttention(ED(input_feat_dim=64, hidden_feat_dim=64, output_feat_dim=64, mode="test"))
    st = torch.load(os.path.join(ref_shim.REFERENCE_ROOT, "weights", "feat_aggregation_3d.pth"), map_location="cpu",
                    weights_only=False)
    model.load_state_dict(st["state_dict"] if "state_dict" in st else st)
    model.eval()
    for k, v in model.state_dict().items():
        out["agg/" + k] = v.numpy()
    V, n_pts = 3000, 3600
    feats = rng.normal(size=(V, 64)).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    coords = rng.uniform(0, 50, size=(V, 3)).astype(np.float32)
    p2v = rng.integers(0, V, n_pts)
    p2v[rng.random(n_pts) < 0.1] = -1                                                    # points without a voxel
    psz = [1, 2, 13, 64, 200, 7, 576, 31, 900]
    order = rng.permutation(n_pts)
    spt, o = [], 0
    for n in psz:
        spt.append(torch.from_numpy(np.sort(order[o:o + n])))
        o += n
    with torch.no_grad():
        f, c = model.aggregation(spt, torch.from_numpy(feats)[None], torch.from_numpy(coords)[None], torch.from_numpy(p2v))
    out.update(agg_feats=feats, agg_coords=coords, agg_p2v=p2v, agg_spt_ptr=np.concatenate([[0], np.cumsum(psz)]).astype(np.int32),
               agg_spt_idx=np.concatenate([x.numpy() for x in spt]), agg_out_feat=f.numpy(), agg_out_coord=c.numpy())
    np.savez_compressed(os.path.join(GOLD, "nets_shipped.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    makers = dict(rigid=make_rigid, f2s3_filter=make_f2s3_filter, knn=make_knn, desc=make_desc,
                  rigidity=make_rigidity, dips=make_dips, lifting=make_lifting,
                  coarse=make_coarse, merge=make_merge, nets=make_nets)
    for name in (sys.argv[1:] or list(makers)):          # `python -m oracle.make_golden dips` refreshes one file
        makers[name]()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
