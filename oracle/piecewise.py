"""TEST INFRASTRUCTURE -- fp64 restatement of the reference's `Piecewise_ICP(cfg)` (rows G1, A5, F5).

Follows src/piecewise_icp.py:89-216.  The octree is Open3D 0.19.0
`geometry.Octree` (third-party, not vendored, not installed, no reference test) -> the octree
part is PARITY UNPINNED and restated from the published algorithm of
open3d/geometry/Octree.cpp:
  ConvertFromPointCloud(size_expand=0): center=(min+max)/2, h=max(center-min),
      origin=min(min_bound, center-h), size=2h; every point inserted by recursive descent
  IsPointInBound: origin <= p < origin+size  (points on the max faces are dropped)
  child index = x + 2y + 4z (x = p.x >= origin.x + size/2 ...), recursion to exactly max_depth
  Traverse: pre-order DFS, children 0..7; a True return on an internal node skips its subtree
Everything else (centroids, 1-NN between centroid sets, mean+std threshold, per-cell
translation, output row order) follows the reference file directly.
"""
import numpy as np
from scipy.spatial import cKDTree

INTERNAL_MIN_POINTS = 250   # piecewise_icp.py:52 (hard-coded early stop, quirk q9)


def bbox_corners(lo, hi):
    """Open3D AxisAlignedBoundingBox.GetBoxPoints order (only membership matters here)."""
    ex = hi - lo
    return np.array([lo, lo + [ex[0], 0, 0], lo + [0, ex[1], 0], lo + [0, 0, ex[2]],
                     hi, lo + [0, ex[1], ex[2]], lo + [ex[0], 0, ex[2]], lo + [ex[0], ex[1], 0]])


def octree_frame(points):
    mn = points.min(0)
    mx = points.max(0)
    center = (mn + mx) / 2
    h = (center - mn).max()
    origin = np.minimum(mn, center - h)
    return origin, 2.0 * h


def octree_codes(points, origin, size, depth):
    """Leaf code per point (base-8 digits x+2y+4z, root digit most significant) and the in-bound
    mask, by the same comparisons as Octree::InsertPointRecurse."""
    p = np.asarray(points, np.float64)
    inb = np.all((origin <= p) & (p < origin + size), axis=1)
    node_origin = np.broadcast_to(origin, p.shape).copy()
    code = np.zeros(p.shape[0], np.int64)
    s = size
    for _ in range(depth):
        s = s / 2.0
        bit = p >= node_origin + s
        node_origin = node_origin + bit * s
        code = code * 8 + bit[:, 0] + 2 * bit[:, 1] + 4 * bit[:, 2]
    return code, inb


def visited_leaf_cells(code, inb, depth, number_points_min):
    """Leaves reached by the DFS with the 250-point early stop, in traversal order.
    Returns (leaf codes ascending == traversal order, counts)."""
    c = code[inb]
    ok = np.ones(c.shape[0], bool)
    for lvl in range(depth):                       # internal nodes are levels 0 .. depth-1
        anc = c >> (3 * (depth - lvl))
        u, inv, cnt = np.unique(anc, return_inverse=True, return_counts=True)
        ok &= cnt[inv] >= INTERNAL_MIN_POINTS      # :52
    u, cnt = np.unique(c, return_counts=True)
    reach = np.isin(u, np.unique(c[ok]))
    sel = reach & (cnt >= number_points_min)       # :55
    return u[sel], cnt[sel]


def cell_centroids(points, code, inb, leaves):
    """fp64 mean of the points of each listed leaf (:58-61)."""
    pos = np.searchsorted(leaves, code)
    pos = np.clip(pos, 0, max(len(leaves) - 1, 0))
    hit = inb & (len(leaves) > 0) & (leaves[pos] == code if len(leaves) else False)
    cent = np.zeros((len(leaves), 3))
    for a in range(3):
        cent[:, a] = np.bincount(pos[hit], weights=points[hit, a], minlength=len(leaves))
    n = np.bincount(pos[hit], minlength=len(leaves))
    return cent / n[:, None], pos, hit


def piecewise_icp(src, tgt, smax, number_points_min):
    """Returns dict(dvfs (N',6), dvfms (N',4), depth, centroids_src, centroids_tgt, nn, dist,
    thr, stable mask, src_cell_of_point, ...).  Raises ValueError like np.vstack([]) does in the
    reference when no cell is unstable (:197)."""
    src = np.asarray(src, np.float64)
    tgt = np.asarray(tgt, np.float64)
    lo = np.minimum(src.min(0), tgt.min(0))                                   # :90-99
    hi = np.maximum(src.max(0), tgt.max(0))
    corners = bbox_corners(lo, hi)
    tgt_a = np.vstack([tgt, corners])                                         # :104
    src_a = np.vstack([src, corners])                                         # :105
    depth = int(np.ceil(np.log2((hi - lo).max() / smax)))                     # :108-109
    depth = max(depth, 0)
    out = dict(depth=depth)
    cells = {}
    for name, pts in (("src", src_a), ("tgt", tgt_a)):
        origin, size = octree_frame(pts)                                      # :115-118
        code, inb = octree_codes(pts, origin, size, depth)
        leaves, cnt = visited_leaf_cells(code, inb, depth, number_points_min)  # :127,:131
        cent, pos, hit = cell_centroids(pts, code, inb, leaves)
        cells[name] = dict(origin=origin, size=size, code=code, inb=inb, leaves=leaves,
                           count=cnt, centroid=cent, pos=pos, hit=hit, pts=pts)
    cs, ct = cells["src"]["centroid"], cells["tgt"]["centroid"]
    _, nn = cKDTree(ct).query(cs, k=1)                                        # :138-148
    dist = np.linalg.norm(cs - ct[nn], axis=1)                                # :152
    thr = dist.mean() + dist.std()                                            # :154-156
    stable = dist <= thr                                                      # :160-161
    S = cells["src"]

    def points_of_cell_at(c):
        """octree_source.locate_leaf_node(centroid)[0].indices (:172,:190)."""
        cc, ib = octree_codes(c[None, :], S["origin"], S["size"], depth)
        if not ib[0]:
            return np.zeros(0, np.int64)
        return np.nonzero(S["inb"] & (S["code"] == cc[0]))[0]

    st_cent = np.unique(cs[stable], axis=0)                                   # :169
    st_idx = [points_of_cell_at(c) for c in st_cent]                          # :171-173
    st_pts = src_a[np.concatenate(st_idx)] if st_idx else np.zeros((0, 3))
    stable_dvfs = np.hstack([st_pts, st_pts])                                 # :175
    un_rows = []
    for c_s, c_t in zip(cs[~stable], ct[nn][~stable]):                        # :185-194
        p = src_a[points_of_cell_at(c_s)]
        un_rows.append(np.hstack([p, p + (c_t - c_s)]))
    unstable_dvfs = np.vstack(un_rows)                                        # :197 (raises if empty)
    dvfs = np.vstack([stable_dvfs, unstable_dvfs])                            # :201-202
    mag = np.linalg.norm(dvfs[:, :3] - dvfs[:, 3:6], axis=1)
    out.update(dvfs=dvfs, dvfms=np.hstack([dvfs[:, :3], mag[:, None]]), centroids_src=cs,
               centroids_tgt=ct, nn=nn, dist=dist, thr=thr, stable=stable,
               leaves_src=S["leaves"], leaves_tgt=cells["tgt"]["leaves"],
               n_stable_pts=stable_dvfs.shape[0], n_src=src_a.shape[0])
    return out
