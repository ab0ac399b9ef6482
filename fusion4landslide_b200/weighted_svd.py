"""Host mirror of the reference's scripts/weighted_svd.py (rows D1, D2, D4): same signatures, CUDA kernels
underneath (one segment per batch element).

    weighted_svd                         scripts/weighted_svd.py:10-55
    weighted_procrustes                  scripts/weighted_svd.py:58-129
    refine_local_rigid_correspondences   scripts/weighted_svd.py:132-159
"""
import torch

from . import ops
from .functions import _batched_ptr, _dev_f32


def weighted_procrustes(src_points, ref_points, weights=None, weight_thresh=0.0, eps=1e-7, return_transform=True,
                        return_rmse=True):
    """Rigid transform src -> ref by weighted SVD; (B,N,3) or (N,3) inputs.
    return_transform=True gives the (B,)4x4 on the CPU like the reference does (quirk q1,
    scripts/weighted_svd.py:119 builds it with torch.eye(4) and never moves it); otherwise (R, t) on the
    input's device."""
    squeeze_first = src_points.ndim == 2
    s = _dev_f32(src_points)
    r = _dev_f32(ref_points, s.device)
    if squeeze_first:
        s, r = s.unsqueeze(0), r.unsqueeze(0)
    B, N = s.shape[0], s.shape[1]
    w = None
    if weights is not None:
        w = _dev_f32(weights, s.device).reshape(B, N).reshape(-1).contiguous()
    R, t, _ = ops.segmented_kabsch(s.reshape(-1, 3), r.reshape(-1, 3), _batched_ptr(B, N, s.device), w=w, eps=eps,
                                   weight_thresh=weight_thresh, variant=0)
    if return_transform:
        T = torch.eye(4).unsqueeze(0).repeat(B, 1, 1)
        T[:, :3, :3] = R.cpu()
        T[:, :3, 3] = t.cpu()
        return T.squeeze(0) if squeeze_first else T
    if squeeze_first:
        return R.squeeze(0), t.squeeze(0)
    return R, t


def weighted_svd(src_pts, tgt_pts, eps=1e-6, weights=None, weight_thresh=0.0, return_transform=True):
    """Older variant kept for signature completeness (scripts/weighted_svd.py:10-55; unused by the mains):
    same estimator, results on the GPU."""
    R, t = weighted_procrustes(src_pts, tgt_pts, weights=weights, weight_thresh=weight_thresh, eps=eps,
                               return_transform=False)
    if not return_transform:
        return R, t
    T = torch.eye(4, device=R.device).repeat(R.shape[0], 1, 1) if R.ndim == 3 else torch.eye(4, device=R.device)
    T[..., :3, :3] = R
    T[..., :3, 3] = t
    return T


def refine_local_rigid_correspondences(corr_neigh_2, refine_type='SVD', weights=None):
    """D1 + `< 1 m` residual prune + 4x4 on the GPU (scripts/weighted_svd.py:132-159).
    Returns (pruned correspondences (k',6), T (4,4) float32 cuda)."""
    if refine_type != 'SVD':
        raise NotImplementedError("refine_type 'RANSAC' is Open3D ransac_registration and stays in the reference")
    corr = _dev_f32(corr_neigh_2)
    K = corr.shape[0]
    if K == 0:                                                         # empty in, empty out (and the identity)
        return corr.reshape(0, 6), torch.eye(4, device=corr.device)
    src, tgt = corr[:, :3].contiguous(), corr[:, 3:6].contiguous()
    ptr = torch.tensor([0, K], dtype=torch.int32, device=corr.device)
    w = None if weights is None else _dev_f32(weights, corr.device).reshape(-1)
    R, t, _, res = ops.segmented_kabsch(src, tgt, ptr, w=w, eps=1e-6, variant=0, want_res=True)
    keep = res < 1.0                                                   # max_res = 1 (:144-147)
    T = torch.eye(4, device=corr.device)
    T[:3, :3] = R[0]
    T[:3, 3] = t[0]
    return corr[keep], T
