"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the DIPs patch front-end, src/data_loader.py:16-109
(Preprocess_Dataset.extract_patch).  numpy fp64 like the reference.

Pinned against the reference: oracle/make_golden.py runs the UNMODIFIED src/data_loader.py through
oracle/ref_shim.py with a stand-in for the one Open3D class it touches (KDTreeFlann.search_radius_vector_3d,
restated below as `radius_search`: nanoflann semantics -- squared distance summed x, y, z in fp64, strict
`d2 < r2`, sorted by distance) and writes tests/golden/dips_patches.npz; tests/test_oracle_golden.py checks this
module against it.  The KD-tree itself is third-party (open3d==0.19.0, requirements.txt:1): its tie order for
EQUAL distances is implementation-defined -- ties are ordered by index here.
"""
import numpy as np
from scipy.spatial import cKDTree

_EPS = 1e-6     # data_loader.py:6


def radius_search(tree, pts, pt, radius):
    """(idx, d2) of the points with d2 < radius^2, ascending in (d2, idx)."""
    cand = np.asarray(tree.query_ball_point(pt, radius * (1.0 + 1e-9) + 1e-12), dtype=np.int64)
    d = pts[cand] - pt
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    keep = d2 < radius * radius
    cand, d2 = cand[keep], d2[keep]
    order = np.lexsort((cand, d2))
    return cand[order], d2[order]


def extract_all(pt, tree, pts, radius):
    """data_loader.py:37-97 up to (not including) the zero padding and the random choice.
    Returns ptall (n,3) fp64 in distance order, lRg (3,3) or None, idx (n)."""
    idx, d2 = radius_search(tree, pts, pt, radius)
    ptall = pts[idx].T                                            # (3, n)
    if ptall.shape[1] > 10:
        ptnn = pts[idx[1:]].T
        dist = np.sqrt(d2[1:])
        vd = ptnn - pt[:, None]
        cov = 1.0 / ptnn.shape[0] * vd @ vd.T                     # :50 (shape[0] is 3)
        a, v = np.linalg.eigh(cov)                                # symmetric: same eigenvectors as np.linalg.eig
        n_hat = v[:, np.argmin(a)]
        zp = n_hat if np.sum(n_hat @ (-vd)) > 0 else -n_hat       # :58
        proj = vd.T @ zp
        vt = vd - np.outer(zp, proj)
        alpha = (radius - dist) ** 2
        beta = proj ** 2
        acc = vt @ (alpha * beta)
        nrm = np.linalg.norm(acc)
        xp = acc / (nrm + _EPS) if abs(nrm) < _EPS else acc / nrm  # :66-72
        yp = np.cross(xp, zp)
        lRg = np.asarray([xp, yp, zp]).T
        out = (lRg.T @ (ptall - pt[:, None])).T / radius
        return out, lRg, idx
    return ptall.T / radius, None, idx                            # :91-94 (not centred: reference quirk)


def sample(ptall, inds, num_points=256):
    """data_loader.py:98-105 with the random indices given."""
    if ptall.shape[0] < num_points:
        ptall = np.concatenate((ptall, np.zeros((num_points - ptall.shape[0], 3))))
    return ptall[inds]


def patches(queries, ref, radius, inds, num_points=256):
    """(n,3,num_points) f32 like Preprocess_Dataset.__getitem__, counts (n), lrf (n,9) rows xp, yp, zp."""
    tree = cKDTree(ref)
    out = np.zeros((len(queries), 3, num_points), np.float32)
    cnt = np.zeros(len(queries), np.int32)
    lrf = np.zeros((len(queries), 9))
    for i, pt in enumerate(queries):
        pa, lRg, idx = extract_all(pt, tree, ref, radius)
        cnt[i] = idx.size
        if lRg is not None:
            lrf[i] = lRg.T.ravel()
        out[i] = sample(pa, inds[i], num_points).T.astype(np.float32)
    return out, cnt, lrf
