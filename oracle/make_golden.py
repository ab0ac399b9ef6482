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
Open3D ICP / Octree, hnswlib and faiss cannot be run anywhere here -> no golden, parity unpinned.
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


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    makers = dict(rigid=make_rigid, f2s3_filter=make_f2s3_filter, knn=make_knn, desc=make_desc,
                  rigidity=make_rigidity, dips=make_dips, lifting=make_lifting)
    for name in (sys.argv[1:] or list(makers)):          # `python -m oracle.make_golden dips` refreshes one file
        makers[name]()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
