"""Host mirror of the hot-path pieces of src/coarse_to_fine_matching_base.py and
src/coarse_to_fine_matching.py: the reference's free functions keep their names and signatures, the
stage methods of `Coarse2Fine_Base` become functions over the same tensors (SURVEY 9.1 field names).

    refine_dvfs_with_threshold                                   base.py:48-97            (A4)
    merge_correspondences_by_priority_with_distance_threshold    coarse_to_fine_matching.py:40-118 (M1)
    compute_median_resolution                                    base.py:2716-2754        (A1)
    voxel_subsampling_maps                                       base.py:1038-1057        (A2)
    global_matches_from_3d                                       base.py:2756-2923        (B2)
    coarse_matching_3d / coarse_matching_2d                      base.py:2966-3011, 3016-3070 (B3, B4)
    fine_matching_with_different_types                           base.py:3236-3457        (F2 F3 D2 E1 D5 A4)
"""
import torch

from . import ops, pipeline
from .functions import _dev_f32

I32 = torch.int32


def refine_dvfs_with_threshold(src_pts, transformed_src_pts, tgt_pts, distance_threshold=0.1, batch_size=1024):
    """[src | nearest tgt of the transformed src] for matches with d^2 < distance_threshold^2 (strict),
    (K,6).  `batch_size` is accepted for signature compatibility (one launch handles everything)."""
    s = _dev_f32(src_pts)
    ts = _dev_f32(transformed_src_pts, s.device)
    t = _dev_f32(tgt_pts, s.device)
    if s.shape[0] == 0 or t.shape[0] == 0:
        return torch.empty((0, 6), device=s.device)
    qp = torch.tensor([0, ts.shape[0]], dtype=I32, device=s.device)
    rp = torch.tensor([0, t.shape[0]], dtype=I32, device=s.device)
    thr = torch.tensor([distance_threshold], dtype=torch.float32, device=s.device)
    nn, _ = ops.segmented_nn(ts, t, qp, rp, thr=thr)
    keep = nn >= 0
    return torch.cat([s[keep], t[nn[keep].long()]], dim=1)


def merge_correspondences_by_priority_with_distance_threshold(corres_list, distance_threshold=1e-3, search_type="faiss"):
    """Keep level 0 entirely; from each later level drop the rows whose source xyz has an already kept point
    within `distance_threshold` (squared-distance test D < thr^2).  Every `search_type` of the reference maps
    to the same exact grid search (its default 'faiss' HNSW index is approximate)."""
    if search_type not in ("faiss", "kdtree", "cdist"):
        raise ValueError("unknown search_type %r" % (search_type,))
    if not corres_list:
        raise IndexError("corres_list is empty")
    merged = [corres_list[0]]
    pool = _dev_f32(corres_list[0][:, :3])
    thr2 = float(distance_threshold) ** 2
    for level in range(1, len(corres_list)):
        cur = corres_list[level]
        xyz = _dev_f32(cur[:, :3], pool.device)
        if xyz.shape[0] == 0:
            merged.append(cur)
            continue
        if pool.shape[0] == 0:
            valid = torch.ones(xyz.shape[0], dtype=torch.bool, device=pool.device)
        else:
            _, d2 = ops.knn_grid(xyz, pool, 1)
            valid = ~(d2[:, 0] < thr2)
        merged.append(cur[valid.to(cur.device)])
        pool = torch.cat([pool, xyz[valid]], dim=0).contiguous()
    return torch.cat(merged, dim=0)


def compute_median_resolution(src_pts, tgt_pts):
    """max over the epochs of the median distance to the nearest other point (k=2 self query); device scalar."""
    return ops.median_resolution(_dev_f32(src_pts), _dev_f32(tgt_pts))


def _pixels_xyz0(p, device):
    """(n,2) pixel coordinates -> (n,3) f32 device points with z = 0 (the grid kNN does not bin a flat axis)."""
    import numpy as np
    if isinstance(p, np.ndarray):
        p = torch.from_numpy(np.ascontiguousarray(p))
    p = p.to(device=device, dtype=torch.float32)
    return torch.cat([p[:, :2], torch.zeros((p.shape[0], 1), dtype=torch.float32, device=device)], 1).contiguous()


def map_corr_2d_to_3d(corres_2d, src_pixel, tgt_pixel, pixel_thres, reverse=False):
    """`map_corr_2d_to_3d` (base.py:387-427; rgb_guided.py:590-639): lift image matches to projected 3D points.
    Every projected source point takes its nearest match in the source image (2-D), follows it to the target image
    and takes the nearest projected target point there; both hops must be shorter than `pixel_thres` pixels.
    corres_2d (K,4) [u_src, v_src, u_tgt, v_tgt]; src_pixel (N,2), tgt_pixel (M,2).  Returns device tensors
    (index into tgt_pixel (N,) i64, mask (N,) bool, matched rows of corres_2d (N,4) f64).  Two exact grid-kNN
    launches replace the two cKDTree builds + queries; pixel coordinates are compared in f32.
    reverse=True is `map_corr_2d_to_3d_tgt2src` (base.py:431-472)."""
    import numpy as np
    dev = src_pixel.device if (torch.is_tensor(src_pixel) and src_pixel.is_cuda) else torch.device("cuda:0")
    c = torch.from_numpy(np.ascontiguousarray(corres_2d, dtype=np.float64)) if isinstance(corres_2d, np.ndarray) else corres_2d
    c = c.to(device=dev, dtype=torch.float64)
    a, b = (tgt_pixel, src_pixel) if reverse else (src_pixel, tgt_pixel)
    ca, cb = (c[:, 2:4], c[:, :2]) if reverse else (c[:, :2], c[:, 2:4])
    i1, d1 = ops.knn_grid(_pixels_xyz0(a, dev), _pixels_xyz0(ca, dev), 1)
    rows = c[i1[:, 0].long()]
    hop = (rows[:, :2] if reverse else rows[:, 2:4])
    i2, d2 = ops.knn_grid(_pixels_xyz0(hop, dev), _pixels_xyz0(b, dev), 1)
    thr2 = float(pixel_thres) ** 2
    mask = (d1[:, 0] < thr2) & (d2[:, 0] < thr2)
    return i2[:, 0].long(), mask, rows


def map_corr_2d_to_3d_tgt2src(corres_2d, src_pixel, tgt_pixel, pixel_thres):
    """base.py:431-472: the same lifting from the target image to the source image."""
    return map_corr_2d_to_3d(corres_2d, src_pixel, tgt_pixel, pixel_thres, reverse=True)


def voxel_down_sample(pts, voxel_size):
    """Open3D `pcd.voxel_down_sample(voxel_size)` (base.py:1024-1025) for an (n,3) array / tensor / Open3D cloud:
    fp64 voxel means, rows in ascending voxel order (Open3D's order is its hash map's; same rows)."""
    import numpy as np
    if hasattr(pts, "points"):
        pts = np.asarray(pts.points)
    if isinstance(pts, np.ndarray):
        pts = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float64))
    dev = pts.device if pts.is_cuda else torch.device("cuda:0")
    return ops.voxel_downsample(pts.to(device=dev, dtype=torch.float64).contiguous(), voxel_size)


def voxel_subsampling(src_pts, tgt_pts):
    """`Coarse2Fine_Base._voxel_subsampling` (base.py:1012-1057) on tensors: adaptive voxel size = median resolution,
    voxel means of both epochs, nearest raw point of every voxel and the inverse maps.  Returns a dict with the
    reference's field names."""
    voxel = float(compute_median_resolution(src_pts, tgt_pts).item())
    out = {"voxel_size": voxel}
    for name, p in (("src", src_pts), ("tgt", tgt_pts)):
        sub = voxel_down_sample(p, voxel).float()
        v2p, p2v = voxel_subsampling_maps(sub, p)
        out[name + "_pts_sub"] = sub
        out["idx_voxel2pts_" + name] = v2p
        out["idx_pts2voxel_" + name] = p2v
    return out


def voxel_subsampling_maps(pts_sub, pts_raw):
    """idx_voxel2pts (N_sub,) = nearest raw point of every voxel centroid, idx_pts2voxel (N,) = inverse map with
    -1 default (base.py:1038-1057; duplicate targets: the largest voxel index wins, like a sequential scatter)."""
    sub = _dev_f32(pts_sub)
    raw = _dev_f32(pts_raw, sub.device)
    idx, _ = ops.knn_grid(sub, raw, 1)
    v2p = idx[:, 0].long()
    p2v = torch.full((raw.shape[0],), -1, dtype=torch.int64, device=sub.device)
    p2v.scatter_reduce_(0, v2p, torch.arange(sub.shape[0], device=sub.device), reduce="amax", include_self=True)
    return v2p, p2v


def global_matches_from_3d(feat_src, feat_tgt, src_pts_sub, tgt_pts_sub, idx_voxel2pts_src, idx_voxel2pts_tgt,
                           n_src_raw, max_magnitude, algo="auto"):
    """Exact descriptor 1-NN src->tgt, magnitude gate, scatter to raw indices: corres_3d_voxel_from_3d_idx
    (N_raw,2) int64 with -1 = none.  Also returns the labels (N_sub,) int32."""
    fs = _dev_f32(feat_src)
    ft = _dev_f32(feat_tgt, fs.device)
    labels, _ = ops.desc_nn(fs, ft, algo=algo)
    corres = ops.scatter_global_matches(labels, _dev_f32(src_pts_sub, fs.device), _dev_f32(tgt_pts_sub, fs.device),
                                        idx_voxel2pts_src.to(fs.device, torch.int64).contiguous(),
                                        idx_voxel2pts_tgt.to(fs.device, torch.int64).contiguous(), max_magnitude,
                                        int(n_src_raw))
    return corres, labels


def coarse_matching_3d(spt_coord_src, spt_feat_src, spt_coord_tgt, spt_feat_tgt, max_magnitude,
                       coarse_refinement_3d_type="nn_mutual"):
    """Feature-space NN between superpoints under the coordinate gate; 'nn_mutual' keeps mutual pairs only.
    Returns (src patch positions m, tgt patch positions j*(m)) int64."""
    fs = _dev_f32(spt_feat_src)
    dev = fs.device
    ri, _, ci, _ = ops.desc_nn(fs, _dev_f32(spt_feat_tgt, dev), a_xyz=_dev_f32(spt_coord_src, dev),
                               b_xyz=_dev_f32(spt_coord_tgt, dev), max_mag=max_magnitude, both_dirs=True, algo="exact")
    mask = ri >= 0
    if coarse_refinement_3d_type == "nn_mutual":
        m_of_j = ci[ri.clamp(min=0).long()]
        mask &= m_of_j == torch.arange(ri.shape[0], device=dev, dtype=ci.dtype)
    elif coarse_refinement_3d_type != "only_max_mag":
        raise ValueError("unknown coarse_refinement_3d_type %r" % (coarse_refinement_3d_type,))
    m = torch.nonzero(mask).reshape(-1)
    return m, ri[m].long()


def coarse_matching_2d(corres_3d_from_2d_idx, sp_idx, sp_ptr, idx_pts2spt_tgt, idx_spt_tgt):
    """2D-vote coarse matching: for every source patch the target patch most of its 2D-lifted matches fall
    into.  Returns (src patch positions, tgt patch positions in idx_spt_tgt, tie flags).
    idx_spt_tgt is ascending (prepare_pts2spt_dict): the winning label is located in it by binary search on the device,
    so no label table has to be sized from a device-side maximum (one host synchronisation: the size of the result)."""
    dev = corres_3d_from_2d_idx.device
    lab_t = idx_pts2spt_tgt.to(dev, I32).contiguous()
    spt_t = idx_spt_tgt.to(dev).long().contiguous()
    best, cnt, flag = ops.vote_tgt_patch(corres_3d_from_2d_idx.contiguous(), sp_idx, sp_ptr, lab_t, None)   # raw labels
    if spt_t.numel() == 0:
        e = torch.zeros(0, dtype=torch.int64, device=dev)
        return e, e, torch.zeros(0, dtype=torch.bool, device=dev)
    lab = best.long()
    pos = torch.searchsorted(spt_t, lab.clamp(min=0)).clamp(max=spt_t.numel() - 1)
    local = torch.where((lab >= 0) & (spt_t[pos] == lab), pos, torch.full_like(pos, -1))        # base.py:3062-3064
    m = torch.nonzero(local >= 0).reshape(-1)
    if bool((flag == 255).any().item()):
        raise RuntimeError("coarse_matching_2d: a source patch votes for more than 512 distinct target patches")
    return m, local[m], flag[m] == 1


def fine_matching_with_different_types(src_pts, tgt_pts, label_src, label_tgt, corres_3d, corres_2d=None, pairs=None,
                                       config=None, median_resolution=None, num_min_matches_for_small_patch=10):
    """The per-patch-pair loop of base.py:3254-3438 as one launch sequence.  `pairs` = (src patch positions,
    tgt patch positions) from the coarse stage (default: equal labels).  Returns (FineResult, TileInputs)."""
    tile = pipeline.prepare_tile(_dev_f32(src_pts), _dev_f32(tgt_pts), label_src, label_tgt, corres_3d, corres_2d,
                                 min_pts=num_min_matches_for_small_patch, pairs=pairs)
    cfg = config or pipeline.FineConfig()
    if median_resolution is None:
        r, _ = pipeline.displacement_field(tile, cfg)
    else:
        r = ops.fine_matching(tile.src, tile.tgt, tile.sp_idx, tile.sp_ptr, tile.tp_idx, tile.tp_ptr,
                              tile.tgt_patch_of_point, tile.pair_tgt_patch, corr3d=tile.corr3d, corr2d=tile.corr2d,
                              d_median_resolution=median_resolution, n_src_items=tile.n_src_items,
                              n_tgt_items=tile.n_tgt_items, **cfg.fine_kwargs())
    return r, tile
