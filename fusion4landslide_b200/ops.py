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
