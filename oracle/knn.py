"""TEST INFRASTRUCTURE -- exact xyz nearest-neighbour oracle (fp64).

Follows the reference's exact-kNN call sites (rows A1-A5 of SURVEY 8a):
  src/coarse_to_fine_matching_base.py:2716-2754 / src/f2s3.py:481-508   k=2 self-kNN, median of
        the 2nd distance, max over the two epochs          (sklearn NearestNeighbors kd_tree)
  src/coarse_to_fine_matching_base.py:1038-1042            1-NN voxel -> raw point (scipy cKDTree)
  src/functions.py:127-144                                 compute_c2c (sklearn kd_tree, k=1)
  src/coarse_to_fine_matching_base.py:48-97                refine_dvfs_with_threshold
        (Open3D KDTreeFlann.search_knn_vector_3d(pt, 1), keep iff d2 < thr^2)
  src/piecewise_icp.py:134-149                             centroid 1-NN (Open3D KDTreeFlann)
All of them are exact searches; scipy's cKDTree in fp64 is the checker.  Tie order of the
third-party trees is implementation-defined, so near-ties are flagged, not compared.
"""
import numpy as np
from scipy.spatial import cKDTree

EPS_XYZ_REL = 1e-6   # documented tie: |d2_a - d2_b| <= EPS_XYZ_REL * max(d2_a, d2_b) (+ 1e-12 absolute)


def knn_exact(q, r, k, workers=-1):
    """Exact kNN.  Returns idx (N,k) int64 and squared distances d2 (N,k) fp64, ascending."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    r = np.ascontiguousarray(r, dtype=np.float64)
    tree = cKDTree(r)
    d, i = tree.query(q, k=k, workers=workers)
    if k == 1:
        d = d[:, None]
        i = i[:, None]
    return i.astype(np.int64), d * d


def knn_bruteforce(q, r, k):
    """O(NM) fp64, lowest index wins exact ties; for small cases and tie analysis."""
    q = np.asarray(q, np.float64)
    r = np.asarray(r, np.float64)
    d2 = ((q[:, None, :] - r[None, :, :]) ** 2).sum(-1)
    idx = np.argsort(d2, axis=1, kind="stable")[:, :k]
    return idx.astype(np.int64), np.take_along_axis(d2, idx, 1)


def tie_rows(q, r, k, workers=-1):
    """Rows whose k-th and (k+1)-th (or any adjacent pair within the first k+1) squared
    distances are closer than the documented epsilon -> index comparison exempt."""
    kk = min(k + 1, r.shape[0])
    _, d2 = knn_exact(q, r, kk, workers)
    if kk < 2:
        return np.zeros(q.shape[0], bool)
    gap = np.diff(d2, axis=1)
    tol = EPS_XYZ_REL * d2[:, 1:] + 1e-12      # relative to the larger distance of each adjacent pair
    return (gap <= tol).any(axis=1)


def median_resolution(src, tgt):
    """base.py:2716-2754 / f2s3.py:481-508: max over epochs of median 2nd-NN distance."""
    out = []
    for p in (src, tgt):
        _, d2 = knn_exact(p, p, 2)
        out.append(np.median(np.sqrt(d2[:, 1])))
    return max(out)


def compute_c2c(source_pc, target_pc):
    """src/functions.py:127-144 -> (n,1) distances."""
    _, d2 = knn_exact(source_pc, target_pc, 1)
    return np.sqrt(d2)


def refine_dvfs_with_threshold(src_pts, transformed_src_pts, tgt_pts, distance_threshold=0.1):
    """base.py:48-97.  Returns (rows (k,6), kept mask, nn idx).  Strict d2 < thr^2 (:82)."""
    src_pts = np.asarray(src_pts)
    tgt_pts = np.asarray(tgt_pts)
    if tgt_pts.shape[0] == 0 or src_pts.shape[0] == 0:
        return np.zeros((0, 6), src_pts.dtype), np.zeros(src_pts.shape[0], bool), np.zeros(src_pts.shape[0], np.int64)
    idx, d2 = knn_exact(transformed_src_pts, tgt_pts, 1, workers=1)
    keep = d2[:, 0] < float(distance_threshold) ** 2
    rows = np.concatenate([src_pts[keep], tgt_pts[idx[keep, 0]]], axis=1)
    return rows, keep, idx[:, 0]
