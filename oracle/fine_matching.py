"""TEST INFRASTRUCTURE -- restatement of the reference's per-patch fine-matching hot loop.

Follows src/coarse_to_fine_matching_base.py:3236-3457 (`fine_matching_with_different_types`),
SURVEY section 9.4, with the helpers it calls:
  :3259-3274  correspondence selection by torch.isin                       (row F2)
  :3299-3328  rigidity / isometry check                                    (row F3)
  :3338-3342  refine_local_rigid_correspondences -> weighted_procrustes    (rows D1/D2)
  :3353-3368  Open3D point-to-point ICP on the MATCHED points, init T_svd  (row E1, unpinned)
  :3371-3405  apply T to all src patch points (+ inverse for tgt2src)      (row D5)
  :3414-3436  assign_then_nn: refine_dvfs_with_threshold, appended twice   (row A4, quirk q4)
fp64 inside the fits (Open3D is fp64; the reference's SVD is fp32), f32 at the same places the
reference casts (`torch.tensor(..., dtype=torch.float32)` at :3366, f32 apply at :3373).
`weighting_svd` is False in every shipped config (quirk q2) and is not restated.
"""
import numpy as np

from . import icp as _icp
from . import knn as _knn
from . import rigid as _rigid


class FineParams:
    def __init__(self, mode="only_3d", remove_low_quality_patch_matches=True,
                 num_min_matches_for_quality_check=10, thres_dist_diff=0.5,
                 thres_inlier_ratio=0.15, num_min_fine_match=10, icp_refine=True,
                 assign_type="assign_then_nn", output_tgt2src=False, icp_threshold=0.1,
                 median_max_resolution=0.1, icp_max_iter=30):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def select_corr(corr_idx, ps, pt):
    """:3259-3261 -- rows of corr_idx[ps] whose tgt index is in pt."""
    rows = corr_idx[ps]
    return rows[np.isin(rows[:, 1], pt)]


def fine_matching(src_pts, tgt_pts, corr3d, corr2d, spt_src, spt_tgt, p: FineParams):
    """Returns dict with dense (list per pair or None), sparse, tgt2src, T (Q,4,4) f32,
    status (Q,) int8: 0 = fitted, 1 = rejected by quality check, 2 = too few matches,
    plus per-pair K, fitness, rmse, ratio_inlier, dist_mean."""
    src_pts = np.asarray(src_pts, np.float32)
    tgt_pts = np.asarray(tgt_pts, np.float32)
    Q = len(spt_src)
    out = dict(dense=[None] * Q, sparse=[None] * Q, tgt2src=[None] * Q,
               T=np.tile(np.eye(4, dtype=np.float32), (Q, 1, 1)),
               T64=np.tile(np.eye(4), (Q, 1, 1)), Tsvd64=np.tile(np.eye(4), (Q, 1, 1)),
               status=np.zeros(Q, np.int8), K=np.zeros(Q, np.int64),
               fitness=np.zeros(Q), rmse=np.zeros(Q), iters=np.zeros(Q, np.int64),
               ratio_inlier=np.zeros(Q), dist_mean=np.zeros(Q), corr=[None] * Q,
               nn_idx=[None] * Q, nn_keep=[None] * Q, thr=np.zeros(Q))
    for i in range(Q):
        ps, pt = np.asarray(spt_src[i]), np.asarray(spt_tgt[i])
        parts = []
        if p.mode in ("only_3d", "fusion"):
            parts.append(select_corr(corr3d, ps, pt))
        if p.mode in ("only_2d", "fusion"):
            parts.append(select_corr(corr2d, ps, pt))
        corr = np.concatenate(parts, axis=0) if len(parts) > 1 else parts[0]     # :3269-3274
        K = corr.shape[0]
        out["K"][i] = K
        out["corr"][i] = corr
        A = src_pts[corr[:, 0]]
        B = tgt_pts[corr[:, 1]]
        if p.remove_low_quality_patch_matches and K >= p.num_min_matches_for_quality_check:
            ratio, dmean = _rigid.rigidity_check(A, B, p.thres_dist_diff)        # :3310-3317
            out["ratio_inlier"][i], out["dist_mean"][i] = ratio, dmean
            if ratio <= p.thres_inlier_ratio or dmean >= p.thres_dist_diff:      # :3320
                out["status"][i] = 1
                continue
        if K < p.num_min_fine_match:                                             # :3338
            out["status"][i] = 2
            continue
        _, Tsvd = _rigid.refine_local_rigid_correspondences(np.concatenate([A, B], 1))   # :3341
        out["Tsvd64"][i] = Tsvd
        if not p.icp_refine:
            continue
        res = _icp.icp_point_to_point(A, B, Tsvd, p.icp_threshold, p.icp_max_iter)       # :3358
        out["fitness"][i], out["rmse"][i], out["iters"][i] = res["fitness"], res["inlier_rmse"], res["iters"]
        T64 = res["transformation"]
        out["T64"][i] = T64
        T = T64.astype(np.float32)                                               # :3363-3366
        out["T"][i] = T
        S = src_pts[ps]
        Tg = tgt_pts[pt]
        moved = (T[:3, :3] @ S.T).T + T[:3, 3]                                   # :3373-3374 (f32)
        out["dense"][i] = np.hstack([S, moved]).astype(np.float32)               # :3404-3405
        if p.output_tgt2src:
            back = (T[:3, :3].T @ (Tg - T[:3, 3]).T).T                           # :3389-3390
            out["tgt2src"][i] = np.hstack([back, Tg]).astype(np.float32)
        if p.assign_type == "assign_all_src":
            movedA = (T[:3, :3] @ A.T).T + T[:3, 3]
            out["sparse"][i] = np.hstack([A, movedA]).astype(np.float32)         # :3412-3413
        elif p.assign_type == "assign_then_nn":
            thr = res["inlier_rmse"] * 2.0                                       # :3420
            if np.isnan(thr) or np.isinf(thr):
                thr = p.median_max_resolution
            thr = max(thr, p.median_max_resolution * 1.0)                        # :3423
            out["thr"][i] = thr
            rows, keep, nn = _knn.refine_dvfs_with_threshold(S, moved, Tg, thr)  # :3427-3429
            out["nn_idx"][i], out["nn_keep"][i] = nn, keep
            out["sparse"][i] = np.vstack([rows, rows]).astype(np.float32)        # :3430,:3436 (q4)
    return out


def stack(parts, width=6):
    parts = [x for x in parts if x is not None]
    return np.vstack(parts) if parts else np.zeros((0, width), np.float32)
