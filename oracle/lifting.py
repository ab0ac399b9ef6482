"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the 2D-match -> 3D-point lifting of
src/coarse_to_fine_matching_base.py:387-472 (map_corr_2d_to_3d, map_corr_2d_to_3d_tgt2src; twins in
src/rgb_guided.py:590-684): every projected source point takes the nearest image match (2-D, k=2 query, first hit
used), follows it to the other image and takes the nearest projected target point there; both hops must be closer
than `pixel_thres` pixels.

Pinned against the reference: oracle/make_golden.py (make_lifting) calls the UNMODIFIED functions through
oracle/ref_shim.py -> tests/golden/map_corr_2d.npz.  scipy's cKDTree is what the reference itself calls.
"""
import numpy as np
from scipy.spatial import cKDTree


def map_corr_2d_to_3d(corres_2d, src_pixel, tgt_pixel, pixel_thres, reverse=False):
    """reverse=False: base.py:387-427; reverse=True: the tgt2src twin, base.py:431-472 (roles of the columns and of
    the two pixel sets swapped).  Returns indices (N,), mask (N,) bool, the matched rows of corres_2d (N,4)."""
    corres_2d = np.asarray(corres_2d)
    a, b = (np.asarray(tgt_pixel), np.asarray(src_pixel)) if reverse else (np.asarray(src_pixel), np.asarray(tgt_pixel))
    ca, cb = (corres_2d[:, 2:4], corres_2d[:, :2]) if reverse else (corres_2d[:, :2], corres_2d[:, 2:4])
    d1, i1 = cKDTree(ca).query(a, k=2)
    rows = corres_2d[i1[:, 0], :]
    d2, i2 = cKDTree(b).query(cb[i1[:, 0]], k=2)
    mask = (d1[:, 0] < pixel_thres) & (d2[:, 0] < pixel_thres)
    return i2[:, 0], mask, rows
