"""TEST INFRASTRUCTURE -- fp64 numpy restatement of the reference's rigid-fit functions.

Follows (reference paths):
  scripts/weighted_svd.py:58-129   weighted_procrustes              (row D1 of SURVEY 8a)
  scripts/weighted_svd.py:132-159  refine_local_rigid_correspondences (D2)
  src/functions.py:12-85           kabsch_transformation_estimation (D3)
  src/functions.py:88-104          transformation_residuals
  src/functions.py:107-124         transform_point_cloud            (D5)
  src/models/outlier_classifier.py:71-105  filter_input tail        (F4)
Pinned by tests/golden/rigid_*.npz, produced from the unmodified reference functions.
"""
import numpy as np


def _svd3(H):
    """torch.svd convention: H = U diag(S) V^T, S descending."""
    U, S, Vt = np.linalg.svd(H)
    return U, S, Vt.T


def weighted_procrustes(src_points, ref_points, weights=None, weight_thresh=0.0, eps=1e-7):
    """scripts/weighted_svd.py:84-129 for one (N,3) pair.  Returns R (3,3), t (3,) in fp64."""
    s = np.asarray(src_points, dtype=np.float64)
    r = np.asarray(ref_points, dtype=np.float64)
    n = s.shape[0]
    w = np.ones(n) if weights is None else np.asarray(weights, dtype=np.float64).copy()
    w = np.where(w < weight_thresh, 0.0, w)                      # :98
    w = w / (w.sum() + eps)                                      # :99
    w = w[:, None]
    cs = (s * w).sum(0, keepdims=True)                           # :102
    cr = (r * w).sum(0, keepdims=True)                           # :103
    H = (s - cs).T @ (w * (r - cr))                              # :107
    U, _, V = _svd3(H)                                           # :111
    d = np.sign(np.linalg.det(V @ U.T))                          # :114  (q7: sign can be 0)
    R = V @ np.diag([1.0, 1.0, d]) @ U.T                         # :115
    t = cr[0] - R @ cs[0]                                        # :117
    return R, t


def procrustes_transform(src_points, ref_points, weights=None, weight_thresh=0.0, eps=1e-7):
    R, t = weighted_procrustes(src_points, ref_points, weights, weight_thresh, eps)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def refine_local_rigid_correspondences(corr, weights=None):
    """scripts/weighted_svd.py:132-151 ('SVD' branch).  corr (K,6) -> (pruned corr, T 4x4)."""
    corr = np.asarray(corr, dtype=np.float64)
    R, t = weighted_procrustes(corr[:, :3], corr[:, 3:6], weights, 0.0, eps=1e-6)   # :134-142
    res = (R @ corr[:, :3].T).T + t - corr[:, 3:6]                                  # :143
    keep = np.linalg.norm(res, axis=1) < 1.0                                        # :145-147
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return corr[keep], T


def transformation_residuals(x1, x2, R, t):
    """src/functions.py:100-104 for one (n,3) pair; t is (3,) or (3,1)."""
    t = np.asarray(t, dtype=np.float64).reshape(3)
    rec = (np.asarray(R, np.float64) @ np.asarray(x1, np.float64).T).T + t
    return np.linalg.norm(rec - np.asarray(x2, np.float64), axis=1)


def kabsch(x1, x2, weights=None, normalize_w=True, eps=1e-7, w_threshold=0.0):
    """src/functions.py:32-85 for b=1.  Returns R (3,3), t (3,), res (n,), flag.

    Quirks kept: weights normalised by (sum+eps) (:36-37) and the means divided AGAIN by
    (sum(w_normalised)+eps) (:51-52); the reflection fix uses the raw determinant (:73-75).
    """
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    n = x1.shape[0]
    w = np.ones(n) if weights is None else np.asarray(weights, dtype=np.float64).copy()
    if normalize_w:
        w = w / (w.sum() + eps)
    if w_threshold > 0:
        w = np.where(w < w_threshold, 0.0, w)
    sw = w.sum() + eps
    m1 = (w[:, None] * x1).sum(0) / sw
    m2 = (w[:, None] * x2).sum(0) / sw
    c1 = x1 - m1
    c2 = x2 - m2
    cov = c1.T @ (w[:, None] * c2)                               # :57-60 (diag_embed made explicit)
    if not np.all(np.isfinite(cov)):
        R = np.eye(3)
        t = np.zeros(3)
        return R, t, transformation_residuals(x1, x2, R, t), True    # :62-71
    U, _, V = _svd3(cov)
    det = np.linalg.det(V.T @ U.T)                               # :73 (det(V^T U^T) == det(V U^T))
    R = V @ np.diag([1.0, 1.0, det]) @ U.T                       # :75-77
    t = m2 - R @ m1                                              # :80
    return R, t, transformation_residuals(x1, x2, R, t), False


def lower_median(x):
    """torch.median semantics: the lower of the two middle elements."""
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    return x[(x.size - 1) // 2]


def filter_input_tail(x1, x2, scores, coeff=1.0):
    """src/models/outlier_classifier.py:71-105 after the network forward.

    scores = relu(tanh(net)) per correspondence.  Returns dict(rot_est, trans_est, residuals,
    robust_estimate, n_inliers, median).
    """
    R, t, res, _ = kabsch(x1, x2, scores)                        # :73-74
    med = lower_median(res)
    inl = np.nonzero(res < coeff * med)[0]                       # :80
    robust = False
    if inl.size >= 5 and med < 0.5:                              # :91
        robust = True
        w = np.zeros_like(res)
        w[inl] = 1.0
        R, t, res, _ = kabsch(x1, x2, w)                         # :96-97
    return dict(rot_est=R, trans_est=t, residuals=res, robust_estimate=robust,
                n_inliers=int(inl.size), median=med)


def transform_point_cloud(x1, R, t):
    """src/functions.py:119-122."""
    return (np.asarray(R, np.float64) @ np.asarray(x1, np.float64).T).T + np.asarray(t, np.float64).reshape(3)


def rigidity_check(A, B, thres_dist_diff):
    """src/coarse_to_fine_matching_base.py:3308-3317 (row F3).  Returns ratio_inlier, dist_mean."""
    A = np.asarray(A, np.float64)
    B = np.asarray(B, np.float64)
    K = A.shape[0]
    dA = np.linalg.norm(A[:, None, :] - A[None, :, :], axis=2)
    dB = np.linalg.norm(B[:, None, :] - B[None, :, :], axis=2)
    diff = np.abs(dA - dB)
    num_ele = K * (K - 1) / 2
    dist_mean = np.triu(diff, 1).sum() / num_ele                 # :3315
    ratio = ((diff <= thres_dist_diff).sum() - K) / (num_ele * 2)    # :3316 (q8)
    return ratio, dist_mean
