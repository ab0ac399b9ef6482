"""TEST INFRASTRUCTURE -- exact descriptor-space nearest neighbour and patch matching oracle.

Follows the reference's EXACT branches (the shipped default back-ends, hnswlib / faiss HNSW, are
approximate third-party indexes that are absent here and not reproducible bit-for-bit by
anything -> parity unpinned for them):
  src/coarse_to_fine_matching_base.py:2783-2815  'cdist_cpu' / 'cdist': torch.cdist + min(dim=1)
  src/coarse_to_fine_matching_base.py:2872-2889  magnitude gate + scatter into (N_raw,2)  (B2)
  src/coarse_to_fine_matching_base.py:2966-2995  coarse superpoint matching, mutual NN     (B3)
  src/coarse_to_fine_matching_base.py:3016-3070  2D-vote coarse matching                  (B4)
  src/f2s3.py:273-285                            F2S3 1-NN src->tgt + (N,6) rows           (B1)
  src/coarse_to_fine_matching.py:40-118          level merge with 1e-3 m dedup             (M1)
fp64 squared-L2, first minimal index wins (torch.min semantics).
"""
import numpy as np

EPS_DESC_ABS = 1e-6   # documented tie: |d2_best - d2_second| <= 1e-6 (unit-norm descriptors)


def desc_nn(a, b, block=2048, return_second=False):
    """argmin_j ||a_i - b_j||^2 in fp64, blocked.  Returns idx (N,), d2 (N,), [second-best d2]."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    bn = (b * b).sum(1)
    idx = np.empty(a.shape[0], np.int64)
    d2 = np.empty(a.shape[0])
    d2b = np.empty(a.shape[0])
    for s in range(0, a.shape[0], block):
        x = a[s:s + block]
        D = (x * x).sum(1)[:, None] + bn[None, :] - 2.0 * (x @ b.T)
        j = D.argmin(1)
        idx[s:s + block] = j
        r = np.arange(x.shape[0])
        d2[s:s + block] = D[r, j]
        if return_second:
            D[r, j] = np.inf
            d2b[s:s + block] = D.min(1)
    d2 = np.maximum(d2, 0.0)
    if return_second:
        return idx, d2, d2b
    return idx, d2


def global_matches_from_3d(src_feat, tgt_feat, src_sub, tgt_sub, idx_voxel2pts_src,
                           idx_voxel2pts_tgt, n_src_raw, max_magnitude):
    """base.py:2791-2889 (exact variant).  Returns corres (N_raw,2) i64, labels, keep.
    Duplicate raw targets of the scatter (quirk q5) are resolved as 'largest voxel index wins',
    the order a sequential scatter gives (torch CPU index_put_)."""
    labels, _ = desc_nn(src_feat, tgt_feat)
    mag = np.linalg.norm(np.asarray(src_sub, np.float32) - np.asarray(tgt_sub, np.float32)[labels], axis=1)
    keep = mag <= max_magnitude                                              # :2875-2876
    C = -np.ones((n_src_raw, 2), np.int64)                                   # :2879-2881
    C[:, 0] = np.arange(n_src_raw)
    C[np.asarray(idx_voxel2pts_src)[keep], 1] = np.asarray(idx_voxel2pts_tgt)[labels[keep]]   # :2883-2885
    return C, labels, keep


def coarse_matching_3d(coord_s, feat_s, coord_t, feat_t, max_magnitude, kind="nn_mutual", block=1024):
    """base.py:2966-2995.  Returns (src patch ids m, tgt patch ids j*(m)) of accepted pairs.
    Row blocks of the (S,T) distance matrices (the reference materialises them whole); first minimum wins, as
    torch.min / np.argmin do."""
    cs = np.asarray(coord_s, np.float64)
    ct = np.asarray(coord_t, np.float64)
    fs = np.asarray(feat_s, np.float64)
    ft = np.asarray(feat_t, np.float64)
    S, T = cs.shape[0], ct.shape[0]
    j = np.zeros(S, np.int64)
    dj = np.full(S, np.inf)
    col_best = np.full(T, np.inf)
    m_of_j = np.zeros(T, np.int64)
    ft2 = (ft * ft).sum(1)
    for lo in range(0, S, block):
        hi = min(lo + block, S)
        Dc = np.sqrt(np.maximum(((cs[lo:hi, None, :] - ct[None, :, :]) ** 2).sum(-1), 0))
        Df = np.sqrt(np.maximum((fs[lo:hi] * fs[lo:hi]).sum(1)[:, None] + ft2[None, :] - 2 * fs[lo:hi] @ ft.T, 0))
        Df[Dc > max_magnitude] = np.inf                                      # :2969
        if T:
            jj = Df.argmin(1)                                                # :2972
            j[lo:hi] = jj
            dj[lo:hi] = Df[np.arange(hi - lo), jj]
            cm = Df.argmin(0)                                                # :2979 (first minimum over ALL rows)
            cv = Df[cm, np.arange(T)]
            upd = cv < col_best
            col_best[upd] = cv[upd]
            m_of_j[upd] = cm[upd] + lo
    in_mag = dj < np.inf                                                     # :2989
    if kind == "only_max_mag":
        mask = in_mag
    else:
        mask = in_mag & (m_of_j[j] == np.arange(S))                          # :2982-2986
    m = np.nonzero(mask)[0]
    return m, j[m]


def coarse_matching_2d_vote(corr2d, idx_pts2spt_tgt, spt2pts_src, idx_spt_tgt):
    """base.py:3016-3070.  For each src patch: the tgt patch label with the most 2D matches
    (ties: torch.argsort(descending) order is unspecified for equal counts -> flagged).
    Returns (src patch list index, tgt patch LOCAL index, tie flag) for accepted pairs."""
    idx_spt_tgt = np.asarray(idx_spt_tgt)
    pos = {int(l): k for k, l in enumerate(idx_spt_tgt)}
    src_ids, tgt_ids, ties = [], [], []
    for m, pts in enumerate(spt2pts_src):
        t = corr2d[np.asarray(pts), 1]
        t = t[t >= 0]                                                        # :3020
        if t.size == 0:
            continue                                                         # :3036-3042
        lab = np.asarray(idx_pts2spt_tgt)[t]
        u, c = np.unique(lab, return_counts=True)
        best = c.max()
        winners = u[c == best]
        w = int(winners[0])
        if w not in pos:                                                     # :3062-3064
            continue
        src_ids.append(m)
        tgt_ids.append(pos[w])
        ties.append(winners.size > 1)
    return np.asarray(src_ids, np.int64), np.asarray(tgt_ids, np.int64), np.asarray(ties, bool)


def f2s3_correspondences(src_xyz, tgt_xyz, src_feat, tgt_feat):
    """src/f2s3.py:281-285 with an exact index: rows [src_xyz | tgt_xyz[label]]."""
    labels, _ = desc_nn(src_feat, tgt_feat)
    return np.concatenate([np.asarray(src_xyz), np.asarray(tgt_xyz)[labels]], axis=1), labels


def merge_by_priority(corres_list, distance_threshold=1e-3):
    """coarse_to_fine_matching.py:40-118, exact semantics of the 'faiss' branch
    (IndexHNSWFlat is approximate; D < thr^2 on squared f32 distances, :96-97)."""
    from scipy.spatial import cKDTree
    kept = [np.asarray(corres_list[0])]
    masks = [np.ones(kept[0].shape[0], bool)]
    for lvl in range(1, len(corres_list)):
        cur = np.asarray(corres_list[lvl])
        pool = np.concatenate([k[:, :3] for k in kept], 0).astype(np.float64)
        if pool.shape[0] == 0 or cur.shape[0] == 0:
            dup = np.zeros(cur.shape[0], bool)
        else:
            d, _ = cKDTree(pool).query(cur[:, :3].astype(np.float64), k=1)
            dup = d * d < distance_threshold ** 2
        kept.append(cur[~dup])
        masks.append(~dup)
    return np.concatenate(kept, 0), masks
