"""Host mirror of the hot-path stages of src/f2s3.py::Deformation_Analyze (rows B1, F4, F1, A3).

    correspondence_searching   src/f2s3.py:248-298    exact 1-NN in descriptor space -> rows [src_xyz | tgt_xyz[label]]
    filter_input_tail          src/models/outlier_classifier.py:71-105   everything after the network forward
    correspondence_pruning     src/f2s3.py:318-441    per-supervoxel tail + magnitude gates
The filtering network forward (PointCN stack) stays in PyTorch as in the reference.
"""
import torch

from . import ops
from .functions import _dev_f32, compute_c2c  # noqa: F401  (compute_c2c is the C2C gap filling, f2s3.py:452-467)

I32 = torch.int32


def correspondence_searching(src_xyz, tgt_xyz, src_feat, tgt_feat, algo="auto"):
    """labels (N,) int64 and correspondences (N,6) [src_xyz | tgt_xyz[labels]] (src/f2s3.py:281-285).
    The reference's hnswlib index is approximate; this is its exact limit."""
    fs = _dev_f32(src_feat)
    labels, _ = ops.desc_nn(fs, _dev_f32(tgt_feat, fs.device), algo=algo)
    s = _dev_f32(src_xyz, fs.device)
    t = _dev_f32(tgt_xyz, fs.device)
    return labels.long(), torch.cat([s, t[labels.long()]], dim=1)


def filter_input_tail(corr, scores, seg_ptr, coeff=1.0):
    """Per supervoxel: Kabsch(scores) -> residuals -> res < coeff*median -> (>= 5 inliers and median < 0.5)
    -> refit with 0/1 weights.  corr (K,6) rows grouped by supervoxel (CSR seg_ptr).  Returns rot_est (Q,3,3),
    trans_est (Q,3), robust_estimate (Q) bool, residuals (K) of the final fit."""
    corr = _dev_f32(corr)
    Q = seg_ptr.numel() - 1
    R = torch.empty((Q, 3, 3), dtype=torch.float32, device=corr.device)
    t = torch.empty((Q, 3), dtype=torch.float32, device=corr.device)
    robust = torch.empty((Q,), dtype=torch.uint8, device=corr.device)
    res = torch.empty((corr.shape[0],), dtype=torch.float32, device=corr.device)
    from ._lib import check, lib, ptr, stream_ptr
    check(lib().f4l_f2s3_prune_tail(ptr(corr), ptr(_dev_f32(scores, corr.device).reshape(-1)), ptr(seg_ptr.to(corr.device, I32)),
                                    None, Q, float(coeff), ptr(R), ptr(t), ptr(robust), ptr(res), None,
                                    stream_ptr(corr.device)), "f4l_f2s3_prune_tail")
    return R, t, robust.bool(), res


def correspondence_pruning(corr, scores, seg_ptr, data_dir="", refine_results=False, max_disp_magnitude=0.0,
                           filter_median_magnitude=False):
    """src/f2s3.py:340-441 after the network: keep mask per row (robust and refine_results -> whole supervoxel,
    else score > 0.99999), saved rows are the UNREFINED coordinates (quirk q6), then magnitude <= max and the
    optional 30 x median gate.  Returns (kept rows (k,6), magnitudes (k,), keep mask (K,))."""
    corr = _dev_f32(corr)
    coeff = 2.5 if 'Rockfall_Simulator' in data_dir else 1.0
    _, _, robust, _ = filter_input_tail(corr, scores, seg_ptr, coeff)
    sc = _dev_f32(scores, corr.device).reshape(-1)
    seg_ptr = seg_ptr.to(corr.device)
    seg_of_row = torch.repeat_interleave(torch.arange(seg_ptr.numel() - 1, device=corr.device),
                                         (seg_ptr[1:] - seg_ptr[:-1]).long())
    keep = sc > 0.99999
    if refine_results:
        keep = keep | robust[seg_of_row]
    rows = corr[keep].contiguous()
    mask, mag = ops.magnitude_mask(rows, max_mag=max_disp_magnitude if max_disp_magnitude > 0 else float("inf"))
    sel = mask.bool()
    rows, mag = rows[sel], mag[sel]
    if filter_median_magnitude and rows.shape[0] > 0:
        n = mag.shape[0]
        kth = ops.select_kth(mag.contiguous(), (n - 1) // 2, n // 2)          # np.median: mean of the middle two
        med = (0.5 * (kth[0] + kth[1])).reshape(1)
        m2 = ops.magnitude_mask(rows.contiguous(), d_max=med, factor=30.0, strict=True, want_mag=False).bool()
        rows, mag = rows[m2], mag[m2]
    return rows, mag, keep
