"""Thin torch-tensor wrappers over the C ABI (include/f4l_b200.h).

torch is plumbing here: it owns device memory (caching allocator) and the current stream; every
computation below happens in libf4l_b200.so kernels.  All tensors must be CUDA + contiguous.
"""
import torch

from . import _lib
from ._lib import F4LError, check, lib, ptr, stream_ptr

I32 = torch.int32
F32 = torch.float32
F64 = torch.float64


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


def segmented_kabsch(src, tgt, seg_start, seg_count=None, w=None, src_idx=None, tgt_idx=None,
                     eps=1e-7, weight_thresh=0.0, variant=0, want_res=False, want_T64=False):
    """K-d.  Returns R (Q,3,3) f32, t (Q,3) f32, flag (Q) u8, [res (K)], [T64 (Q,4,4)]."""
    Q = seg_start.numel() if seg_count is not None else seg_start.numel() - 1
    K = (src_idx if src_idx is not None else src).shape[0]
    R = _empty((Q, 3, 3), F32, src)
    t = _empty((Q, 3), F32, src)
    flag = _empty((Q,), torch.uint8, src)
    res = _empty((K,), F32, src) if want_res else None
    T64 = _empty((Q, 4, 4), F64, src) if want_T64 else None
    check(lib().f4l_segmented_kabsch(
        ptr(src, F32), ptr(tgt, F32), ptr(src_idx, I32, True), ptr(tgt_idx, I32, True),
        ptr(w, F32, True), ptr(seg_start, I32), ptr(seg_count, I32, True), Q, eps, weight_thresh,
        variant, ptr(R), ptr(t), ptr(T64, F64, True), ptr(res, F32, True), ptr(flag),
        stream_ptr(src.device)), "f4l_segmented_kabsch")
    out = [R, t, flag]
    if want_res:
        out.append(res)
    if want_T64:
        out.append(T64)
    return tuple(out)


def apply_transforms(pts, seg_start, T, seg_count=None, idx=None, out_start=None, seg_skip=None,
                     inverse=False, n_rows=None, want_mag=True):
    """K-f.  Returns dvf (rows,6) f32 and mag (rows) f32."""
    Q = seg_start.numel() if seg_count is not None else seg_start.numel() - 1
    if n_rows is None:
        n_rows = (idx if idx is not None else pts).shape[0]
    dvf = _empty((n_rows, 6), F32, pts)
    mag = _empty((n_rows,), F32, pts) if want_mag else None
    T = T.reshape(Q, 16)
    check(lib().f4l_apply_transforms(
        ptr(pts, F32), ptr(idx, I32, True), ptr(seg_start, I32), ptr(seg_count, I32, True),
        ptr(out_start, I32, True), ptr(seg_skip, torch.uint8, True), Q, ptr(T, F32), int(inverse),
        ptr(dvf), ptr(mag, F32, True), stream_ptr(pts.device)), "f4l_apply_transforms")
    return dvf, mag


def rigidity_check(src, tgt, seg_start, thres_dist_diff, seg_count=None, src_idx=None, tgt_idx=None):
    """K-c.  Returns ratio_inlier (Q) f32, dist_mean (Q) f32."""
    Q = seg_start.numel() if seg_count is not None else seg_start.numel() - 1
    ratio = _empty((Q,), F32, src)
    dmean = _empty((Q,), F32, src)
    check(lib().f4l_rigidity_check(
        ptr(src, F32), ptr(tgt, F32), ptr(src_idx, I32, True), ptr(tgt_idx, I32, True),
        ptr(seg_start, I32), ptr(seg_count, I32, True), Q, float(thres_dist_diff), ptr(ratio),
        ptr(dmean), stream_ptr(src.device)), "f4l_rigidity_check")
    return ratio, dmean


def segmented_median(x, seg_start, seg_count=None):
    """torch.median (lower median) of each segment of x (K) f32 -> (Q) f32."""
    Q = seg_start.numel() if seg_count is not None else seg_start.numel() - 1
    med = _empty((Q,), F32, x)
    check(lib().f4l_segmented_median(ptr(x, F32), ptr(seg_start, I32), ptr(seg_count, I32, True), Q, ptr(med),
                                     stream_ptr(x.device)), "f4l_segmented_median")
    return med


_WS = {}


def _workspace(nbytes, device):
    """Grow-only per-device scratch buffer (torch caching allocator owns the memory)."""
    if device.type != "cuda":
        raise _lib.F4LError("tensor must live on a CUDA device (no CPU fallback)")
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def knn_grid(q, r, k, max_radius=0.0, cell=0.0):
    """K-a.  Exact kNN of q (N,3) among r (M,3).  Returns idx (N,k) i32, d2 (N,k) f32."""
    N, M = q.shape[0], r.shape[0]
    idx = _empty((N, k), I32, q)
    d2 = _empty((N, k), F32, q)
    nbytes = lib().f4l_knn_grid_workspace_bytes(N, M)
    ws = _workspace(nbytes, q.device)
    check(lib().f4l_knn_grid(ptr(q, F32), N, ptr(r, F32), M, k, float(max_radius), float(cell),
                             ptr(idx), ptr(d2), ptr(ws), ws.numel(), stream_ptr(q.device)),
          "f4l_knn_grid")
    return idx, d2


def knn_ties(q, r, k, eps_rel=1e-6, max_radius=0.0, cell=0.0):
    """Tie flags (N) u8 of knn_grid(q, r, k): rows whose neighbour order is decided by a distance difference below
    eps_rel (relative) -- exempt from index comparison with another exact search (f4l_knn_grid_ties)."""
    N, M = q.shape[0], r.shape[0]
    tie = _empty((N,), torch.uint8, q)
    ws = _workspace(lib().f4l_knn_grid_workspace_bytes(N, M), q.device)
    check(lib().f4l_knn_grid_ties(ptr(q, F32), N, ptr(r, F32), M, int(k), float(max_radius), float(cell), float(eps_rel),
                                  ptr(tie), ptr(ws), ws.numel(), stream_ptr(q.device)), "f4l_knn_grid_ties")
    return tie


def patch_icp(src, tgt, s_start, t_start, s_count=None, t_count=None, src_idx=None, tgt_idx=None,
              T0=None, max_corr_dist=0.1, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6, seg_skip=None,
              want_corr=False, want_fragile=False, tie_eps=None):
    """K-e.  Returns T (Q,4,4) f64, fitness (Q) f64, rmse (Q) f64, iters (Q) i32, [corr (items) i32],
    [fragile (Q) u8: OR of FRAGILE_NN / FRAGILE_INLIER / FRAGILE_STOP -- a decision of the loop was within tie_eps
    (default 1e-9, relative) of flipping, include/f4l_b200.h f4l_patch_icp_ex]."""
    Q = s_start.numel() if s_count is not None else s_start.numel() - 1
    T = _empty((Q, 4, 4), F64, src)
    fit = _empty((Q,), F64, src)
    rmse = _empty((Q,), F64, src)
    iters = _empty((Q,), I32, src)
    n_items = (src_idx if src_idx is not None else src).shape[0]
    corr = _empty((n_items,), I32, src) if want_corr else None
    if T0 is not None:
        T0 = T0.reshape(Q, 16)
    fragile = _empty((Q,), torch.uint8, src) if want_fragile else None
    check(lib().f4l_patch_icp_ex(
        ptr(src, F32), ptr(src_idx, I32, True), ptr(s_start, I32), ptr(s_count, I32, True),
        ptr(tgt, F32), ptr(tgt_idx, I32, True), ptr(t_start, I32), ptr(t_count, I32, True),
        ptr(seg_skip, torch.uint8, True), Q, ptr(T0, F64, True), float(max_corr_dist), int(max_iter),
        float(rel_fitness), float(rel_rmse), ptr(T), ptr(fit), ptr(rmse), ptr(iters),
        ptr(corr, I32, True), ptr(fragile, torch.uint8, True), float(tie_eps or 0.0), stream_ptr(src.device)),
        "f4l_patch_icp_ex")
    out = (T, fit, rmse, iters)
    if want_corr:
        out = out + (corr,)
    if want_fragile:
        out = out + (fragile,)
    return out


FRAGILE_NN, FRAGILE_INLIER, FRAGILE_STOP = 1, 2, 4


def segmented_nn(qpts, rpts, q_start, r_start, q_count=None, r_count=None, qidx=None, ridx=None,
                 T=None, thr=None, want_d2=True):
    """A4.  Returns nn (items) i32 (position inside the reference segment or -1), d2 (items) f32."""
    Q = q_start.numel() if q_count is not None else q_start.numel() - 1
    n_items = (qidx if qidx is not None else qpts).shape[0]
    nn = _empty((n_items,), I32, qpts)
    d2 = _empty((n_items,), F32, qpts) if want_d2 else None
    if T is not None:
        T = T.reshape(Q, 16)
    check(lib().f4l_segmented_nn(
        ptr(qpts, F32), ptr(qidx, I32, True), ptr(q_start, I32), ptr(q_count, I32, True),
        ptr(rpts, F32), ptr(ridx, I32, True), ptr(r_start, I32), ptr(r_count, I32, True), Q,
        ptr(T, F32, True), ptr(thr, F32, True), ptr(nn), ptr(d2, F32, True),
        stream_ptr(qpts.device)), "f4l_segmented_nn")
    return nn, d2


def median_resolution(src, tgt, out=None):
    """A1.  Device scalar (1,) f32: max over epochs of the median nearest-other-point distance."""
    if out is None:
        out = _empty((1,), F32, src)
    nbytes = lib().f4l_median_resolution_workspace_bytes(src.shape[0], tgt.shape[0])
    ws = _workspace(nbytes, src.device)
    check(lib().f4l_median_resolution(ptr(src, F32), src.shape[0], ptr(tgt, F32), tgt.shape[0], ptr(out, F32),
                                      ptr(ws), ws.numel(), stream_ptr(src.device)), "f4l_median_resolution")
    return out


def select_kth(x, k, k2=-1, stride=1, offset=0):
    n = x.numel() // stride
    out = _empty((2,), F32, x)
    ws = _workspace(lib().f4l_select_kth_workspace_bytes(n), x.device)
    check(lib().f4l_select_kth(ptr(x, F32), n, stride, offset, k, k2, ptr(out), ptr(ws), ws.numel(),
                               stream_ptr(x.device)), "f4l_select_kth")
    return out


_DESC_ALGO = {"auto": 0, "tensor": 1, "exact": 2}


def desc_nn(a, b, a_xyz=None, b_xyz=None, max_mag=0.0, both_dirs=False, algo="auto", tie_eps=None):
    """K-b.  Exact descriptor-space nearest neighbour of every row of a (N,D) among b (M,D), D in {32,64}.
    Returns row_idx (N) i32, row_d2 (N) f32 [, col_idx (M), col_d2 (M) when both_dirs]
    [, row_tie (N) u8 [, col_tie (M) u8] when tie_eps is given: another row within tie_eps of the minimum]."""
    N, D = a.shape
    M = b.shape[0]
    row_idx = _empty((N,), I32, a)
    row_d2 = _empty((N,), F32, a)
    col_idx = _empty((M,), I32, a) if both_dirs else None
    col_d2 = _empty((M,), F32, a) if both_dirs else None
    ties = tie_eps is not None
    row_tie = _empty((N,), torch.uint8, a) if ties else None
    col_tie = _empty((M,), torch.uint8, a) if (ties and both_dirs) else None
    ws = _workspace(lib().f4l_desc_nn_workspace_bytes(N, M, D, int(both_dirs)), a.device)
    check(lib().f4l_desc_nn_ex(ptr(a, F32), N, ptr(b, F32), M, D, ptr(a_xyz, F32, True), ptr(b_xyz, F32, True),
                               float(max_mag), int(both_dirs), _DESC_ALGO[algo], ptr(row_idx), ptr(row_d2),
                               ptr(col_idx, I32, True), ptr(col_d2, F32, True), ptr(row_tie, torch.uint8, True),
                               ptr(col_tie, torch.uint8, True), float(tie_eps or 0.0), ptr(ws), ws.numel(),
                               stream_ptr(a.device)), "f4l_desc_nn")
    out = (row_idx, row_d2) + ((col_idx, col_d2) if both_dirs else ())
    if ties:
        out = out + ((row_tie, col_tie) if both_dirs else (row_tie,))
    return out


def scatter_global_matches(labels, src_sub, tgt_sub, voxel2pts_src, voxel2pts_tgt, max_magnitude, n_raw):
    """base.py:2872-2889.  Returns corres (n_raw,2) i64."""
    corres = _empty((n_raw, 2), torch.int64, src_sub)
    ws = _workspace(lib().f4l_scatter_global_matches_workspace_bytes(n_raw), src_sub.device)
    check(lib().f4l_scatter_global_matches(ptr(labels, I32), ptr(src_sub, F32), ptr(tgt_sub, F32), labels.shape[0],
                                           ptr(voxel2pts_src, torch.int64), ptr(voxel2pts_tgt, torch.int64),
                                           float(max_magnitude), ptr(corres), n_raw, ptr(ws), ws.numel(),
                                           stream_ptr(src_sub.device)), "f4l_scatter_global_matches")
    return corres


def labels_to_csr(labels, min_pts=10):
    """prepare_pts2spt_dict (base.py:1301-1351).  labels (N) i64 -> patch_label (P) i64, ptr (P+1) i32, idx (items) i32,
    patch_of_point (N) i32.  Reads the two sizes back (this step precedes the path)."""
    n = labels.shape[0]
    dev = labels.device
    lab = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    p = torch.empty((n + 1,), dtype=I32, device=dev)
    idx = torch.empty((max(n, 1),), dtype=I32, device=dev)
    pop = torch.empty((max(n, 1),), dtype=I32, device=dev)
    counts = torch.empty((2,), dtype=I32, device=dev)
    ws = _workspace(lib().f4l_labels_to_csr_workspace_bytes(n), dev)
    check(lib().f4l_labels_to_csr(ptr(labels, torch.int64), n, int(min_pts), ptr(lab), ptr(p), ptr(idx), ptr(pop), ptr(counts),
                                  ptr(ws), ws.numel(), stream_ptr(dev)), "f4l_labels_to_csr")
    P_, items = counts.tolist()
    return lab[:P_], p[:P_ + 1], idx[:items], pop[:n]


def gather_pairs_csr(p, idx, sel):
    """Concatenated point lists of the selected patches: (out_ptr (Q+1) i32, out_idx i32, total items)."""
    Q = sel.numel()
    dev = p.device
    sel = sel.to(I32).contiguous()
    out_ptr = torch.empty((Q + 1,), dtype=I32, device=dev)
    ws = _workspace(lib().f4l_gather_pairs_csr_workspace_bytes(Q), dev)
    check(lib().f4l_gather_pairs_csr(ptr(p, I32), ptr(idx, I32), ptr(sel, I32), Q, ptr(out_ptr), None, 0, ptr(ws), ws.numel(),
                                     stream_ptr(dev)), "f4l_gather_pairs_csr")
    total = int(out_ptr[-1].item()) if Q else 0
    out_idx = torch.empty((max(total, 1),), dtype=I32, device=dev)
    if total:
        check(lib().f4l_gather_pairs_csr(ptr(p, I32), ptr(idx, I32), ptr(sel, I32), Q, ptr(out_ptr), ptr(out_idx), total,
                                         ptr(ws), ws.numel(), stream_ptr(dev)), "f4l_gather_pairs_csr")
    return out_ptr, out_idx[:total], total


def vote_tgt_patch(corr2d, sp_idx, sp_ptr, label_tgt, label_to_local=None):
    """B4.  Returns best (P) i32, best_count (P) i32, flag (P) u8."""
    Pn = sp_ptr.numel() - 1
    best = _empty((Pn,), I32, corr2d)
    cnt = _empty((Pn,), I32, corr2d)
    flag = _empty((Pn,), torch.uint8, corr2d)
    check(lib().f4l_vote_tgt_patch(ptr(corr2d, torch.int64), ptr(sp_idx, I32), ptr(sp_ptr, I32), Pn, ptr(label_tgt, I32),
                                   label_tgt.numel(), ptr(label_to_local, I32, True),
                                   0 if label_to_local is None else label_to_local.numel(), ptr(best), ptr(cnt),
                                   ptr(flag), stream_ptr(corr2d.device)), "f4l_vote_tgt_patch")
    return best, cnt, flag


def magnitude_mask(rows, max_mag=0.0, d_max=None, factor=1.0, strict=False, want_mag=True):
    """F1.  rows (K,>=6) f32.  Returns mask (K) u8 [, mag (K) f32]."""
    K = rows.shape[0]
    mask = _empty((K,), torch.uint8, rows)
    mag = _empty((K,), F32, rows) if want_mag else None
    check(lib().f4l_magnitude_mask(ptr(rows, F32), K, rows.shape[1], float(max_mag), ptr(d_max, F32, True), float(factor),
                                   int(strict), ptr(mag, F32, True), ptr(mask), stream_ptr(rows.device)),
          "f4l_magnitude_mask")
    return (mask, mag) if want_mag else mask


def piecewise_icp(src64, tgt64, smax, number_points_min, internal_min_points=250, want_tables=False):
    """K-g.  The reference's Piecewise_ICP on device tensors (f64 (n,3)).  Returns dvfs ((n_src+8),6) f64,
    mag (n_src+8) f64 (both upper bounds), counts (6) i32 [rows, stable rows, Cs, Ct, depth, unstable cells],
    thr (1) f64 [, cent_src, cent_tgt, nn]."""
    n_s, n_t = src64.shape[0], tgt64.shape[0]
    dvfs = _empty((n_s + 8, 6), F64, src64)
    mag = _empty((n_s + 8,), F64, src64)
    counts = _empty((6,), I32, src64)
    thr = _empty((1,), F64, src64)
    cs = _empty((n_s + 8, 3), F64, src64) if want_tables else None
    ct = _empty((n_t + 8, 3), F64, src64) if want_tables else None
    nn = _empty((n_s + 8,), I32, src64) if want_tables else None
    ws = _workspace(lib().f4l_piecewise_icp_workspace_bytes(n_s, n_t), src64.device)
    check(lib().f4l_piecewise_icp(ptr(src64, F64), n_s, ptr(tgt64, F64), n_t, float(smax), int(number_points_min),
                                  int(internal_min_points), ptr(dvfs), ptr(mag), ptr(counts), ptr(thr),
                                  ptr(cs, F64, True), ptr(ct, F64, True), ptr(nn, I32, True), ptr(ws), ws.numel(),
                                  stream_ptr(src64.device)), "f4l_piecewise_icp")
    if want_tables:
        return dvfs, mag, counts, thr, cs, ct, nn
    return dvfs, mag, counts, thr


class FineResult:
    """Outputs of the fused fine-matching stage (device tensors; row counts in `counts`)."""
    __slots__ = ("T", "T64", "status", "K", "fitness", "rmse", "iters", "ratio_inlier", "dist_mean",
                 "dense", "sparse", "tgt2src", "counts", "sparse_pair_rows", "icp_fragile")

    def rows(self):
        """Host sync: slice the row buffers to their true lengths (dense, sparse, tgt2src)."""
        c = self.counts.tolist()
        return (self.dense[:c[0]], self.sparse[:c[1]],
                self.tgt2src[:c[2]] if self.tgt2src is not None else None)


_MODES = {"only_3d": 0, "only_2d": 1, "fusion": 2}
_ASSIGN = {"assign_all_src": 0, "assign_then_nn": 1, "assign_then_nn_once": 2}


class FineCall:
    """One tile's fused fine-matching stage, prepared but not yet enqueued: result tensors, parameter / buffer structs and
    the tile's own workspace.  `run(phases)` enqueues the selected phases on the current stream."""
    __slots__ = ("result", "prm", "bf", "ws", "device", "_keep")

    def run(self, phases=0):
        import ctypes
        self.bf.phases = int(phases)
        check(lib().f4l_fine_matching(ctypes.byref(self.prm), ctypes.byref(self.bf), ptr(self.ws), self.ws.numel(),
                                      stream_ptr(self.device)), "f4l_fine_matching")
        return self.result


PHASE_SELECT, PHASE_FIT_SMALL, PHASE_FIT_LARGE, PHASE_FINISH = 1, 2, 4, 8


def fine_prepare(src_pts, tgt_pts, sp_idx, sp_ptr, tp_idx, tp_ptr, tgt_patch_of_point, pair_tgt_patch,
                 corr3d=None, corr2d=None, mode="only_3d", remove_low_quality_patch_matches=True,
                 num_min_matches_for_quality_check=10, thres_dist_diff=0.5, thres_inlier_ratio=0.15,
                 num_min_fine_match=10, icp_refine=True, assign_type="assign_then_nn",
                 output_tgt2src=False, icp_threshold=0.1, median_max_resolution=0.1,
                 d_median_resolution=None, icp_max_iter=30, n_src_items=None, n_tgt_items=None, out=None,
                 peer_dense=None, median_event=None, own_workspace=False, want_fragile=False,
                 corr3d_tgt=None, corr2d_tgt=None):
    """Builds the FineCall of one tile (see fine_matching for the arguments).  own_workspace: allocate a workspace that
    belongs to this call (needed when the phases of several tiles interleave: fine_fit_tiles); default: the per-stream
    cached one."""
    dev = src_pts.device
    Q = sp_ptr.numel() - 1
    if n_src_items is None:
        n_src_items = int(sp_ptr[-1].item()) if Q > 0 else 0
    if n_tgt_items is None:
        n_tgt_items = int(tp_ptr[-1].item()) if Q > 0 else 0
    r = out
    if r is None:
        r = FineResult()
        r.T = torch.empty((Q, 4, 4), dtype=F32, device=dev)
        r.T64 = torch.empty((Q, 4, 4), dtype=F64, device=dev)
        r.status = torch.empty((Q,), dtype=torch.int8, device=dev)
        r.K = torch.empty((Q,), dtype=I32, device=dev)
        r.fitness = torch.empty((Q,), dtype=F64, device=dev)
        r.rmse = torch.empty((Q,), dtype=F64, device=dev)
        r.iters = torch.empty((Q,), dtype=I32, device=dev)
        r.ratio_inlier = torch.empty((Q,), dtype=F32, device=dev)
        r.dist_mean = torch.empty((Q,), dtype=F32, device=dev)
        r.dense = torch.empty((n_src_items, 6), dtype=F32, device=dev)
        r.sparse = torch.empty((2 * n_src_items, 6), dtype=F32, device=dev)
        r.tgt2src = torch.empty((n_tgt_items, 6), dtype=F32, device=dev) if output_tgt2src else None
        r.counts = torch.empty((4,), dtype=I32, device=dev)
        r.sparse_pair_rows = torch.empty((Q,), dtype=I32, device=dev) if assign_type == "assign_then_nn_once" else None
        r.icp_fragile = torch.empty((Q,), dtype=torch.uint8, device=dev) if want_fragile else None
    prm = _lib.FineParams(_MODES[mode], int(remove_low_quality_patch_matches), int(num_min_matches_for_quality_check),
                          float(thres_dist_diff), float(thres_inlier_ratio), int(num_min_fine_match), int(icp_refine),
                          _ASSIGN[assign_type], int(output_tgt2src), float(icp_threshold),
                          float(median_max_resolution), int(icp_max_iter))
    I64 = torch.int64
    bf = _lib.FineBuffers(
        ptr(src_pts, F32), src_pts.shape[0], ptr(tgt_pts, F32), tgt_pts.shape[0],
        ptr(corr3d, I64, True), ptr(corr2d, I64, True),
        ptr(sp_idx, I32), ptr(sp_ptr, I32), ptr(tp_idx, I32), ptr(tp_ptr, I32),
        ptr(tgt_patch_of_point, I32), ptr(pair_tgt_patch, I32), Q, n_src_items, n_tgt_items,
        ptr(d_median_resolution, F32, True),
        ptr(r.T), ptr(r.T64), ptr(r.status), ptr(r.K), ptr(r.fitness), ptr(r.rmse), ptr(r.iters),
        ptr(r.ratio_inlier), ptr(r.dist_mean), ptr(r.dense), ptr(r.sparse),
        ptr(r.tgt2src, F32, True), ptr(r.counts))
    if median_event is not None:
        bf.median_ready_event = int(median_event.cuda_event)      # d_median_resolution comes from another stream
    spr = getattr(r, "sparse_pair_rows", None)
    if spr is not None:
        bf.sparse_pair_rows = ptr(spr, I32)
    frag = getattr(r, "icp_fragile", None)
    if frag is not None:
        bf.icp_fragile = ptr(frag, torch.uint8)
    # column 1 of the correspondence tables as int32 (n_src): used when the int64 (n_src,2) table is not handed in
    if corr3d is None and corr3d_tgt is not None:
        bf.corr3d_tgt = ptr(corr3d_tgt, I32)
    if corr2d is None and corr2d_tgt is not None:
        bf.corr2d_tgt = ptr(corr2d_tgt, I32)
    if peer_dense:
        if len(peer_dense) > _lib.MAX_PEERS:
            raise _lib.F4LError("at most %d peers" % _lib.MAX_PEERS)
        bf.n_peers = len(peer_dense)
        for i, pp in enumerate(peer_dense):
            bf.peer_dense[i] = int(pp)
    nbytes = lib().f4l_fine_matching_workspace_bytes(n_src_items, n_tgt_items, Q, _MODES[mode])
    c = FineCall()
    c.result, c.prm, c.bf, c.device = r, prm, bf, dev
    c.ws = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=dev) if own_workspace else _workspace(nbytes, dev)
    c._keep = (src_pts, tgt_pts, sp_idx, sp_ptr, tp_idx, tp_ptr, tgt_patch_of_point, pair_tgt_patch, corr3d, corr2d,
               d_median_resolution, median_event, corr3d_tgt, corr2d_tgt)
    return c


def fine_matching(*args, **kw):
    """Fused fine-matching stage of one tile (f4l_fine_matching).  n_*_items = sp_ptr[-1], tp_ptr[-1]
    (pass them to avoid a device->host read).  peer_dense: device pointers (ints) of the slot of `out.dense`
    in every peer GPU's exchange buffer (exchange.PeerExchange); the D5 kernel stores each dense row there too."""
    return fine_prepare(*args, **kw).run(0)


_FIT_QUEUES = {}


def fine_fit_tiles(calls, ctas_per_sm=0):
    """The small-pair fits (rigidity check, Procrustes, ICP; <= 224 matches) of many prepared tiles in ONE persistent
    launch on the current stream (f4l_fine_fit_tiles): run every call's PHASE_SELECT before (stream-ordered), and
    PHASE_FIT_LARGE | PHASE_FINISH after.  The calls must own their workspaces (fine_prepare(own_workspace=True))."""
    import ctypes
    if not calls:
        return
    dev = calls[0].device
    key = (dev.index, stream_ptr(dev))
    q = _FIT_QUEUES.get(key)
    if q is None:
        q = _FIT_QUEUES[key] = torch.zeros((1,), dtype=I32, device=dev)
    for lo in range(0, len(calls), 128):
        chunk = calls[lo:lo + 128]
        arr = (_lib.FineBuffers * len(chunk))(*[c.bf for c in chunk])
        wss = (ctypes.c_void_p * len(chunk))(*[ptr(c.ws) for c in chunk])
        check(lib().f4l_fine_fit_tiles(ctypes.byref(chunk[0].prm), arr, wss, len(chunk), int(ctas_per_sm), ptr(q), stream_ptr(dev)),
              "f4l_fine_fit_tiles")


def peer_push(rows, d_count, peer_ptrs, n_ctas=0):
    """Copy rows[:d_count[0]] (device scalar count) into the same slot of every peer's field (f4l_peer_push) on the
    current stream.  rows: contiguous (n, c) tensor; peer_ptrs: device pointers (ints) of the slot in each peer."""
    import ctypes
    if not peer_ptrs:
        return
    arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    check(lib().f4l_peer_push(ptr(rows), ptr(d_count, I32), int(rows.shape[1] * rows.element_size()), int(rows.shape[0]),
                              arr, len(peer_ptrs), int(n_ctas), stream_ptr(rows.device)), "f4l_peer_push")


class DipsIndex:
    """Ball-query index of one reference cloud (the counterpart of o3d.geometry.KDTreeFlann(pcd),
    data_loader.py:26): the cloud binned for queries of `radius`, resident in a workspace tensor."""

    def __init__(self, ref64, radius):
        if ref64.dtype != F64:
            raise F4LError("DipsIndex: reference cloud must be float64 (the reference works on Open3D's doubles)")
        self.ref = ref64.contiguous()
        self.n_ref = int(ref64.shape[0])
        self.radius = float(radius)
        nbytes = lib().f4l_dips_workspace_bytes(self.n_ref)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=ref64.device)
        check(lib().f4l_dips_build(ptr(self.ref, F64), self.n_ref, self.radius, ptr(self.ws), nbytes,
                                   stream_ptr(ref64.device)), "f4l_dips_build")


def dips_patches(index, query64, num_points=256, ranks=None, seed=0, want_lrf=False, out=None, large=False):
    """K-h.  (n,3,num_points) f32 patches of data_loader.py:37-105 for the rows of query64 (n,3) f64.
    ranks: optional (n,num_points) i32 distance ranks to keep (the reference's `inds`).
    large: the variant that keeps up to 8192 neighbours per query on chip (default: 1408; a query beyond the limit gets a
    zero patch, count tells).  Returns patches, count (n) i32 [, lrf (n,9) f64]."""
    n = int(query64.shape[0])
    patches = out if out is not None else _empty((n, 3, num_points), F32, query64)
    count = _empty((n,), I32, query64)
    lrf = _empty((n, 9), F64, query64) if want_lrf else None
    if large:
        if ranks is not None:
            raise F4LError("dips_patches: the large variant has no ranked mode")
        check(lib().f4l_dips_patches_large(ptr(query64, F64), n, index.n_ref, index.radius, int(num_points),
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(patches, F32), ptr(lrf, F64, True),
                                           ptr(count, I32), ptr(index.ws), index.ws.numel(), stream_ptr(query64.device)),
              "f4l_dips_patches_large")
    else:
        check(lib().f4l_dips_patches(ptr(query64, F64), n, index.n_ref, index.radius, int(num_points),
                                     ptr(ranks, I32, True), int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(patches, F32),
                                     ptr(lrf, F64, True), ptr(count, I32), ptr(index.ws), index.ws.numel(),
                                     stream_ptr(query64.device)), "f4l_dips_patches")
    if want_lrf:
        return patches, count, lrf
    return patches, count


def host_pack_corr_targets(corr, out=None, n_threads=2):
    """HOST tensors: column 1 of an (n,2) int64 correspondence table as int32 (n) -- what the fused stage reads of it
    (`corr3d_tgt` / `corr2d_tgt`); 4 instead of 16 bytes per source point cross PCIe.  `out`: a pinned (n) int32 buffer."""
    if corr.is_cuda or (out is not None and out.is_cuda):
        raise F4LError("host_pack_corr_targets works on host tensors")
    if corr.dtype != torch.int64 or corr.dim() != 2 or corr.shape[1] != 2 or not corr.is_contiguous():
        raise F4LError("corr must be a contiguous (n,2) int64 tensor")
    n = int(corr.shape[0])
    if out is None:
        out = torch.empty((n,), dtype=I32)
    if out.dtype != I32 or out.numel() < n or not out.is_contiguous():
        raise F4LError("out must be a contiguous int32 tensor of at least n elements")
    lib().f4l_host_pack_corr_targets(corr.data_ptr(), n, out.data_ptr(), int(n_threads))
    return out[:n]


def host_expand_sparse(once, pair_rows, out, n_threads=4):
    """HOST tensors: restore the reference's per-pair doubled sparse layout (base.py:3430,3436) from rows emitted
    once (assign_type "assign_then_nn_once").  once (R,6) f32, pair_rows (Q) i32, out (>= 2R,6) f32.  Returns rows."""
    if once.is_cuda or pair_rows.is_cuda or out.is_cuda:
        raise F4LError("host_expand_sparse works on host tensors")
    return int(lib().f4l_host_expand_sparse(once.data_ptr(), pair_rows.data_ptr(), int(pair_rows.numel()),
                                            out.data_ptr(), int(n_threads)))


def voxel_downsample(pts64, voxel_size, want_map=False):
    """Open3D voxel_down_sample (base.py:1024-1025) on the GPU: (n,3) f64 -> centroids (V,3) f64 in ascending voxel
    order [, voxel_of_point (n) i32].  One host read (the voxel count sizes the returned view)."""
    if pts64.dtype != F64:
        raise F4LError("voxel_downsample: points must be float64 (Open3D works on doubles)")
    n = int(pts64.shape[0])
    cent = _empty((max(n, 1), 3), F64, pts64)
    vop = _empty((n,), I32, pts64) if want_map else None
    counts = _empty((1,), I32, pts64)
    nbytes = lib().f4l_voxel_downsample_workspace_bytes(n)
    ws = _workspace(nbytes, pts64.device)
    check(lib().f4l_voxel_downsample(ptr(pts64, F64) if n else None, n, float(voxel_size), ptr(cent, F64),
                                     ptr(vop, I32, True), ptr(counts, I32), ptr(ws), ws.numel(),
                                     stream_ptr(pts64.device)), "f4l_voxel_downsample")
    v = int(counts.item())
    if v < 0:
        raise F4LError("voxel_downsample: the cloud spans more than 2^21 voxels along an axis")
    return (cent[:v], vop) if want_map else cent[:v]


# ---- 8(f) rank 2: per-segment parts of the filtering network / superpoint attention pooling -------------------
def segment_scale_maxabs(x, seg_ptr):
    """rows (K,C) of every segment divided by the segment's max |value| (src/f2s3.py:343)."""
    if x.dtype not in (F32, F64):
        raise F4LError("segment_scale_maxabs: float32 or float64 rows")
    out = torch.empty(x.shape, dtype=F32, device=x.device)
    Q = seg_ptr.numel() - 1
    check(lib().f4l_segment_scale_maxabs(ptr(x), int(x.dtype == F64), ptr(seg_ptr, I32), Q, x.shape[1], ptr(out),
                                         stream_ptr(x.device)), "f4l_segment_scale_maxabs")
    return out


def segment_norm2_relu(y, seg_ptr, eps=1e-3, residual=None):
    """InstanceNorm2d -> BatchNorm2d(batch stats) -> ReLU [-> + residual] per segment and channel (PointCN)."""
    y = y.contiguous()
    out = torch.empty_like(y)
    Q = seg_ptr.numel() - 1
    check(lib().f4l_segment_norm2_relu(ptr(y, F32), ptr(seg_ptr, I32), Q, y.shape[1], float(eps),
                                       ptr(residual, F32, True), ptr(out), stream_ptr(y.device)), "f4l_segment_norm2_relu")
    return out


def segment_attention_pool(Qm, Km, Vm, seg_ptr, scale, max_seg_rows=None):
    """(P,hidden): mean over the rows of each segment of softmax(Q K^T * scale) V.  max_seg_rows: the longest segment
    when the caller knows it (else read back here: one small device->host copy); 0 forces the CUDA-core kernel."""
    Pn = seg_ptr.numel() - 1
    out = _empty((Pn, Qm.shape[1]), F32, Qm)
    if max_seg_rows is None:
        max_seg_rows = int((seg_ptr[1:] - seg_ptr[:-1]).max().item()) if Pn > 0 else 0
    check(lib().f4l_segment_attention_pool(ptr(Qm, F32), ptr(Km, F32), ptr(Vm, F32), ptr(seg_ptr, I32), Pn, Qm.shape[1],
                                           float(scale), int(max_seg_rows), ptr(out), stream_ptr(Qm.device)),
          "f4l_segment_attention_pool")
    return out


def segment_mean(x, seg_ptr):
    Pn = seg_ptr.numel() - 1
    out = _empty((Pn, x.shape[1]), F32, x)
    check(lib().f4l_segment_mean(ptr(x, F32), ptr(seg_ptr, I32), Pn, x.shape[1], ptr(out), stream_ptr(x.device)),
          "f4l_segment_mean")
    return out
