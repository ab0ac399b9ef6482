"""Thin torch-tensor wrappers over the C ABI (include/f4l_b200.h).

torch is plumbing here: it owns device memory (caching allocator) and the current stream; every
computation below happens in libf4l_b200.so kernels.  All tensors must be CUDA + contiguous.
"""
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

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


_WS = {}


def _workspace(nbytes, device):
    """Grow-only per-device scratch buffer (torch caching allocator owns the memory)."""
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


def patch_icp(src, tgt, s_start, t_start, s_count=None, t_count=None, src_idx=None, tgt_idx=None,
              T0=None, max_corr_dist=0.1, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6, seg_skip=None,
              want_corr=False):
    """K-e.  Returns T (Q,4,4) f64, fitness (Q) f64, rmse (Q) f64, iters (Q) i32, [corr (items) i32]."""
    Q = s_start.numel() if s_count is not None else s_start.numel() - 1
    T = _empty((Q, 4, 4), F64, src)
    fit = _empty((Q,), F64, src)
    rmse = _empty((Q,), F64, src)
    iters = _empty((Q,), I32, src)
    n_items = (src_idx if src_idx is not None else src).shape[0]
    corr = _empty((n_items,), I32, src) if want_corr else None
    if T0 is not None:
        T0 = T0.reshape(Q, 16)
    check(lib().f4l_patch_icp(
        ptr(src, F32), ptr(src_idx, I32, True), ptr(s_start, I32), ptr(s_count, I32, True),
        ptr(tgt, F32), ptr(tgt_idx, I32, True), ptr(t_start, I32), ptr(t_count, I32, True),
        ptr(seg_skip, torch.uint8, True), Q, ptr(T0, F64, True), float(max_corr_dist), int(max_iter),
        float(rel_fitness), float(rel_rmse), ptr(T), ptr(fit), ptr(rmse), ptr(iters),
        ptr(corr, I32, True), stream_ptr(src.device)), "f4l_patch_icp")
    out = (T, fit, rmse, iters)
    return out + (corr,) if want_corr else out


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
