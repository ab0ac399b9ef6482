"""Host mirror of the reference's src/functions.py (rows D3, D5, A3 of SURVEY 8a): same names, argument
order, defaults and return shapes; the arithmetic runs in libf4l_b200.so kernels (no CPU fallback).

    kabsch_transformation_estimation  src/functions.py:12-85
    transformation_residuals          src/functions.py:88-104
    transform_point_cloud             src/functions.py:107-124
    compute_c2c                       src/functions.py:127-144
"""
import numpy as np
import torch

from . import ops


def _dev_f32(x, device=None):
    """numpy / CPU tensors are uploaded (the reference accepts numpy here); CUDA tensors pass through."""
    if not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x))
    if not x.is_cuda:
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("fusion4landslide_b200 needs a CUDA device (no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        x = x.to(device)
    return x.to(torch.float32).contiguous()


def _dev_f32_local(a, b):
    """Two clouds as float32 device tensors in a COMMON LOCAL FRAME: float64 inputs (the reference hands Open3D /
    numpy float64 coordinates to sklearn here, src/f2s3.py:453-454) are shifted by their joint minimum in float64
    before the cast, so georeferenced coordinates (~1e6 m, float32 spacing 0.06-0.25 m) lose nothing that matters to a
    distance.  float32 inputs pass through unchanged."""
    ta = a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))
    tb = b if torch.is_tensor(b) else torch.as_tensor(np.asarray(b))
    if ta.dtype != torch.float64 and tb.dtype != torch.float64:
        fa = _dev_f32(ta)
        return fa, _dev_f32(tb, fa.device)
    if not ta.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("fusion4landslide_b200 needs a CUDA device (no CPU fallback)")
        ta = ta.to(torch.device("cuda", torch.cuda.current_device()))
    tb = tb.to(ta.device)
    ta, tb = ta.to(torch.float64), tb.to(torch.float64)
    if ta.numel() == 0 or tb.numel() == 0:
        return ta.float().contiguous(), tb.float().contiguous()
    piv = torch.minimum(ta.min(0).values, tb.min(0).values)
    return (ta - piv).float().contiguous(), (tb - piv).float().contiguous()


def _batched_ptr(b, n, device):
    return torch.arange(0, (b + 1) * n, n, dtype=torch.int32, device=device)


def kabsch_transformation_estimation(x1, x2, weights=None, normalize_w=True, eps=1e-7, best_k=0, w_threshold=0):
    """Weighted Kabsch, batch of b problems of n pairs (src/functions.py:12-85).
    Returns rot [b,3,3], trans [b,3,1], residuals [b,n], flag (True = degenerate -> identity, :62-71).
    `normalize_w=False` differs from the reference only through the eps regularisation (relative 1e-7):
    the kernel always forms w/(sum w + eps) first; rotation and translation are invariant to that scale."""
    x1 = _dev_f32(x1)
    x2 = _dev_f32(x2, x1.device)
    b, n = x1.shape[0], x1.shape[1]
    if weights is None:
        w = torch.ones((b, n), dtype=torch.float32, device=x1.device)
    else:
        w = _dev_f32(weights, x1.device).reshape(b, n)
    if best_k > 0:
        # the reference picks the best_k largest weights of batch element 0 for the whole batch (:43-47)
        idx = torch.topk(w[0], best_k).indices.sort().values
        x1, x2, w = x1[:, idx].contiguous(), x2[:, idx].contiguous(), w[:, idx].contiguous()
        n = best_k
    if w_threshold > 0:
        wn = w / (w.sum(1, keepdim=True) + eps) if normalize_w else w
        w = torch.where(wn < w_threshold, torch.zeros_like(w), w)
    ptr = _batched_ptr(b, n, x1.device)
    R, t, flag, res = ops.segmented_kabsch(x1.reshape(-1, 3), x2.reshape(-1, 3), ptr, w=w.reshape(-1).contiguous(),
                                           eps=eps, variant=1, want_res=True)
    return R, t.reshape(b, 3, 1), res.reshape(b, n), bool(flag.any().item())


def transformation_residuals(x1, x2, R, t):
    """||R x1 + t - x2|| per point, [b,n] (src/functions.py:88-104)."""
    x1 = _dev_f32(x1)
    x2 = _dev_f32(x2, x1.device)
    b, n = x1.shape[0], x1.shape[1]
    T = torch.zeros((b, 4, 4), dtype=torch.float32, device=x1.device)
    T[:, :3, :3] = _dev_f32(R, x1.device).reshape(b, 3, 3)
    T[:, :3, 3] = _dev_f32(t, x1.device).reshape(b, 3)
    T[:, 3, 3] = 1
    dvf, _ = ops.apply_transforms(x1.reshape(-1, 3), _batched_ptr(b, n, x1.device), T, want_mag=False)
    return torch.linalg.norm(dvf[:, 3:6] - x2.reshape(-1, 3), dim=1).reshape(b, n)


def transform_point_cloud(x1, R, t):
    """(R x1^T + t)^T for [n,3] points; numpy in -> numpy out like the reference (src/functions.py:107-124)."""
    as_numpy = not torch.is_tensor(x1)
    p = _dev_f32(x1)
    T = torch.zeros((1, 4, 4), dtype=torch.float32, device=p.device)
    T[0, :3, :3] = _dev_f32(R, p.device).reshape(3, 3)
    T[0, :3, 3] = _dev_f32(t, p.device).reshape(3)
    T[0, 3, 3] = 1
    ptr = torch.tensor([0, p.shape[0]], dtype=torch.int32, device=p.device)
    dvf, _ = ops.apply_transforms(p, ptr, T, want_mag=False)
    out = dvf[:, 3:6]
    return out.cpu().numpy().astype(np.asarray(x1).dtype) if as_numpy else out.contiguous()


def compute_c2c(source_pc, target_pc):
    """Cloud-to-cloud 1-NN distances, [n,1] (src/functions.py:127-144).  numpy in -> numpy float64 out."""
    as_numpy = not torch.is_tensor(source_pc)
    s, t = _dev_f32_local(source_pc, target_pc)
    _, d2 = ops.knn_grid(s, t, 1)
    d = torch.sqrt(d2.to(torch.float64))
    return d.cpu().numpy() if as_numpy else d
