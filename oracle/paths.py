"""TEST INFRASTRUCTURE / CPU ARM -- whole-tile compositions of the pinned oracle pieces, in the order the reference's
entry points run them.  Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.

  c2f_tile    `Coarse2Fine.implement_c2f_matching` (src/coarse_to_fine_matching.py:201-290): voxel subsampling
              (base.py:1012-1057) -> per-level patch lists (:1301-1351) -> descriptor 1-NN + gate + scatter
              (:2756-2923) -> per level: superpoint attention pooling (:2561-2656), coarse matching (:2925-3157), fine
              matching (:3236-3457) -> level merge (coarse_to_fine_matching.py:40-118).  BASELINE configs C3 / C4.
  f2s3_tile   `Deformation_Analyze.correspondence_searching` + `correspondence_pruning` (src/f2s3.py:248-441) with the
              filtering-network weights handed in.  BASELINE config C2.
Every stage reports its own wall time (`seconds` dict) so a bounded CPU sample can be scaled stage by stage.
"""
import time

import numpy as np
from scipy.spatial import cKDTree

from . import desc_nn as odesc
from . import fine_matching as ofm
from . import knn as oknn
from . import rigid as orig
from . import voxel as ovox


def patch_lists(labels, min_pts):
    """prepare_pts2spt_dict (base.py:1301-1351): labels of the kept patches (count > min_pts) ascending, point lists
    ascending."""
    labels = np.asarray(labels)
    order = np.argsort(labels, kind="stable")
    u, start, cnt = np.unique(labels[order], return_index=True, return_counts=True)
    keep = cnt > min_pts
    return u[keep], [order[s:s + c] for s, c in zip(start[keep], cnt[keep])]


def voxel_subsampling(src, tgt, voxel_size, v2p_given=None):
    """base.py:1012-1057 for a given voxel size (= the median resolution of the raw tile, :1023).
    v2p_given: {'src': idx, 'tgt': idx} replaces the kd-tree answer (a voxel of two points has its centroid
    equidistant from both up to f32 rounding: the caller checks those rows as ties and injects one choice so that
    everything downstream is compared on identical maps)."""
    out = {}
    for name, p in (("src", src), ("tgt", tgt)):
        sub64, _ = ovox.voxel_down_sample(np.asarray(p, np.float64), voxel_size)     # :1024-1025
        sub = sub64.astype(np.float32)                                               # pcd2tensor
        raw = np.asarray(p, np.float32)
        _, v2p = cKDTree(raw.astype(np.float64)).query(sub.astype(np.float64), k=1)  # :1038-1042
        out["idx_voxel2pts_kdtree_" + name] = v2p
        if v2p_given is not None:
            v2p = np.asarray(v2p_given[name], np.int64)
        p2v = np.full(raw.shape[0], -1, np.int64)                                    # :1049-1056
        p2v[v2p] = np.arange(sub.shape[0])                # sequential assignment: the last (largest) voxel wins
        out[name + "_pts_sub"], out["idx_voxel2pts_" + name], out["idx_pts2voxel_" + name] = sub, v2p, p2v
    return out


def attention_pool(w, feats, coords, lists, p2v):
    """ClusterFeatureNetWithAttention.aggregation, mode 'test' (cluster_feature_net_self_attention.py:72-103), fp64.
    w: state_dict as numpy arrays.  Returns (P,D) features, (P,3) centroids."""
    g = lambda k: np.asarray(w[k], np.float64)
    Wq, bq, Wk, bk, Wv, bv = (g("self_attention.query.weight"), g("self_attention.query.bias"),
                              g("self_attention.key.weight"), g("self_attention.key.bias"),
                              g("self_attention.value.weight"), g("self_attention.value.bias"))
    Wf, bf = g("self_attention.fc.weight"), g("self_attention.fc.bias")
    W0, b0, W2, b2 = g("mlp.0.weight"), g("mlp.0.bias"), g("mlp.2.weight"), g("mlp.2.bias")
    F, C = [], []
    feats = np.asarray(feats, np.float64)
    coords = np.asarray(coords, np.float64)
    for pts in lists:
        v = p2v[pts]
        v = v[v >= 0]                                                                # :80-81
        x = feats[v]
        Q, K, V = x @ Wq.T + bq, x @ Wk.T + bk, x @ Wv.T + bv
        s = Q @ K.T / np.sqrt(K.shape[1])
        s = np.exp(s - s.max(1, keepdims=True))
        a = s / s.sum(1, keepdims=True)
        o = (a @ V) @ Wf.T + bf
        h = o.mean(0)                                                                # :92
        h = np.maximum(h @ W0.T + b0, 0) @ W2.T + b2                                 # :97
        F.append(h)
        C.append(coords[v].mean(0))                                                  # :99
    return np.asarray(F, np.float32).reshape(len(lists), -1), np.asarray(C, np.float32).reshape(len(lists), 3)


def c2f_tile(src, tgt, labels_src, labels_tgt, feat_raw_src, feat_raw_tgt, agg_weights, voxel_size, corr2d=None,
             coarse="only_3d", fine="only_3d", max_magnitude=5.0, min_pts=10, median_max_resolution=None,
             max_pairs_per_level=None, fine_params=None, v2p_given=None, corr3d_given=None):
    """The fusion method on one tile.  labels_*: list of per-level label arrays (n,).  coarse/fine: 'only_3d' |
    'fusion' (2D-vote pairs first, 2D-lifted matches appended in the fine stage).  max_pairs_per_level bounds the
    fine-matching sample (CPU arm of bench.py); the merge then covers the sampled pairs only."""
    sec = {}
    t0 = time.perf_counter()
    src = np.asarray(src, np.float32)
    tgt = np.asarray(tgt, np.float32)
    vs = voxel_subsampling(src, tgt, voxel_size, v2p_given)
    sec["voxel_subsampling"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    med = oknn.median_resolution(vs["src_pts_sub"], vs["tgt_pts_sub"]) if median_max_resolution is None \
        else median_max_resolution                                                   # base.py:1983 / :2668
    sec["median_resolution"] = time.perf_counter() - t0
    fs = np.asarray(feat_raw_src, np.float32)[vs["idx_voxel2pts_src"]]
    ft = np.asarray(feat_raw_tgt, np.float32)[vs["idx_voxel2pts_tgt"]]
    t0 = time.perf_counter()
    if corr3d_given is not None:            # CPU arm of bench.py: the O(N^2) search is timed on a row sample elsewhere
        corr3d, labels = np.asarray(corr3d_given), None
    else:
        corr3d, labels, _ = odesc.global_matches_from_3d(fs, ft, vs["src_pts_sub"], vs["tgt_pts_sub"], vs["idx_voxel2pts_src"],
                                                         vs["idx_voxel2pts_tgt"], src.shape[0], max_magnitude)
    sec["global_matches_from_3d"] = time.perf_counter() - t0
    prm = fine_params or ofm.FineParams(mode=fine, median_max_resolution=float(med))
    levels = []
    sec["pooling"] = sec["coarse"] = sec["fine"] = 0.0
    for ls, lt in zip(labels_src, labels_tgt):
        lab_s, spt_s = patch_lists(ls, min_pts)
        lab_t, spt_t = patch_lists(lt, min_pts)
        t0 = time.perf_counter()
        f_s, c_s = attention_pool(agg_weights, fs, vs["src_pts_sub"], spt_s, vs["idx_pts2voxel_src"])
        f_t, c_t = attention_pool(agg_weights, ft, vs["tgt_pts_sub"], spt_t, vs["idx_pts2voxel_tgt"])
        sec["pooling"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        m, j = odesc.coarse_matching_3d(c_s, f_s, c_t, f_t, max_magnitude, "nn_mutual")
        tie = np.zeros(0, bool)
        if coarse == "fusion":
            m2, j2, tie = odesc.coarse_matching_2d_vote(corr2d, lt, spt_s, lab_t)
            m, j = np.concatenate([m2, m]), np.concatenate([j2, j])                  # base.py:3139-3146
        sec["coarse"] += time.perf_counter() - t0
        if max_pairs_per_level is not None:
            m, j = m[:max_pairs_per_level], j[:max_pairs_per_level]
        t0 = time.perf_counter()
        o = ofm.fine_matching(src, tgt, corr3d, corr2d, [spt_s[a] for a in m], [spt_t[b] for b in j], prm)
        sec["fine"] += time.perf_counter() - t0
        levels.append(dict(m=m, j=j, tie=tie, fine=o, spt_feat_src=f_s, spt_coord_src=c_s, spt_feat_tgt=f_t,
                           spt_coord_tgt=c_t, n_src_points=int(sum(len(spt_s[a]) for a in m))))
    t0 = time.perf_counter()
    dense, _ = odesc.merge_by_priority([ofm.stack(l["fine"]["dense"]) for l in levels])
    sparse, _ = odesc.merge_by_priority([ofm.stack(l["fine"]["sparse"]) for l in levels])
    sec["merge"] = time.perf_counter() - t0
    return dict(vs, corr3d=corr3d, labels=labels, median_max_resolution=med, levels=levels, dense=dense, sparse=sparse,
                seconds=sec)


def f2s3_tile(src, tgt, feat_src, feat_tgt, svl_labels, weights, coeff=1.0, refine_results=False,
              max_disp_magnitude=5.0, mutual=False, min_pts=10, max_rows=None, max_segments=None, labels_given=None):
    """src/f2s3.py:248-441 with the filtering-network output (`weights`, one per source point) handed in.
    mutual: correspondences whose target's nearest source is another point get weight 0 (BASELINE config C2
    "mutual-NN"; the reference's F2S3 is one-directional).  max_rows / max_segments bound the CPU sample: the
    descriptor search then covers the first max_rows source rows, the pruning the first max_segments supervoxels."""
    sec = {}
    src64, tgt64 = np.asarray(src, np.float64), np.asarray(tgt, np.float64)
    fs, ft = np.asarray(feat_src, np.float32), np.asarray(feat_tgt, np.float32)
    n = src64.shape[0] if max_rows is None else min(max_rows, src64.shape[0])
    t0 = time.perf_counter()
    if labels_given is not None:            # CPU arm of bench.py: the O(N^2) search is timed on a row sample elsewhere
        labels = np.asarray(labels_given)[:n]
    else:
        labels, _ = odesc.desc_nn(fs[:n], ft)                                        # :273-281 (exact index)
    sec["desc_nn_rows"] = n
    sec["desc_nn"] = time.perf_counter() - t0
    back = None
    if mutual:
        t0 = time.perf_counter()
        back, _ = odesc.desc_nn(ft, fs) if max_rows is None else (None, None)
        sec["desc_nn_back"] = time.perf_counter() - t0
    corr = np.concatenate([src64[:n], tgt64[labels]], axis=1)                        # :284-285
    _, lists = patch_lists(np.asarray(svl_labels)[:n] if max_rows is not None else svl_labels, min_pts)
    if max_segments is not None:
        lists = lists[:max_segments]
    w = np.asarray(weights, np.float32)
    t0 = time.perf_counter()
    keep_rows, R, T, robust = [], [], [], []
    for svl in lists:
        X = corr[svl].astype(np.float32)
        sc = w[svl].copy()
        if back is not None:
            sc[back[labels[svl]] != svl] = 0.0
        o = orig.filter_input_tail(X[:, :3], X[:, 3:], sc, coeff)                    # outlier_classifier.py:71-105
        R.append(o["rot_est"]); T.append(o["trans_est"]); robust.append(o["robust_estimate"])
        k = sc > 0.99999                                                             # f2s3.py:363
        if refine_results and o["robust_estimate"]:
            k = np.ones_like(k)                                                      # :351-360
        keep_rows.append(svl[k])
    sec["pruning"] = time.perf_counter() - t0
    sec["pruning_rows"] = int(sum(len(s) for s in lists))
    idx = np.concatenate(keep_rows) if keep_rows else np.zeros(0, np.int64)
    rows = corr[idx]
    mag = np.linalg.norm(rows[:, 3:6] - rows[:, :3], axis=1)
    sel = mag <= max_disp_magnitude                                                  # :392-393 (unconditional)
    rows, mag, idx = rows[sel], mag[sel], idx[sel]
    return dict(labels=labels, back=back, rows=rows, mag=mag, idx=idx, R=np.asarray(R), t=np.asarray(T),
                robust=np.asarray(robust, bool), lists=lists, seconds=sec)


def rgb_local_rigid_refinement(corres_3d_refine, idx_valid_src_refine, segment_patches, icp_thres=0.1, icp_max_iter=30):
    """`Image_DVFs.local_rigid_refinement` (src/rgb_guided.py:981-1062), patch by patch like the reference.
    Returns (mask_valid_local rows, rows [src | T_icp src] stacked, per-patch dict lists)."""
    from . import icp as oicp
    corr_all = np.asarray(corres_3d_refine, np.float32)
    ids = np.asarray(idx_valid_src_refine)
    row_of = {int(v): k for k, v in enumerate(ids)}
    keep_rows, out_rows, per = [], [], []
    for patch in segment_patches:
        idx = np.array([row_of[int(v)] for v in np.asarray(patch).reshape(-1) if int(v) in row_of], np.int64)    # :990-991
        c = corr_all[idx]
        if c.shape[0] == 0:
            per.append(None)
            continue
        R, t = orig.weighted_procrustes(c[:, :3], c[:, 3:6], None, 0.0, eps=1e-6)        # rgb_guided.py:101-109
        res = np.linalg.norm(c[:, :3].astype(np.float64) @ R.T + t - c[:, 3:6], axis=1)
        mask = res < 2.5 * orig.lower_median(res)                                         # :115 torch.median
        keep_rows.append(idx[mask])                                                       # :1001
        T0 = np.eye(4)
        T0[:3, :3], T0[:3, 3] = R, t
        fit0 = oicp.icp_point_to_point(c[:, :3], c[:, 3:6], T0, icp_thres, 0)["fitness"]  # correspondences at the start
        r = oicp.icp_point_to_point(c[:, :3], c[:, 3:6], T0, icp_thres, icp_max_iter)     # :1015-1019
        T = r["transformation"].astype(np.float32)                                        # :1024
        moved = (T[:3, :3] @ c[:, :3].T).T + T[:3, 3]                                     # :1027-1030 (f32)
        out_rows.append(np.hstack([c[:, :3], moved]).astype(np.float32))
        per.append(dict(T0=T0, T=T, iters=r["iters"], fitness=r["fitness"], fitness0=fit0, n=c.shape[0], kept=int(mask.sum())))
    keep = np.concatenate(keep_rows) if keep_rows else np.zeros(0, np.int64)
    rows = np.vstack(out_rows) if out_rows else np.zeros((0, 6), np.float32)
    return keep, rows, per
