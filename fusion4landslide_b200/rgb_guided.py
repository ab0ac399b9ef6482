"""Host mirror of the rigid-refinement helper of src/rgb_guided.py (row F4, RGB-guided variant):

    refine_local_rigid_correspondences   src/rgb_guided.py:99-125
        Procrustes (eps 1e-6) -> residuals -> mask = res < 2.5 * median(res) -> mask_2 = (kept fraction >= 0.70)
`refine_local_rigid_correspondences_batched` is the same over many patches at once (CSR segments): one
launch sequence instead of the reference's per-patch Python loop (rgb_guided.py:981-1062).
"""
import torch

from . import ops
from .functions import _dev_f32


def refine_local_rigid_correspondences_batched(corr, seg_ptr):
    """corr (K,6) rows [src | tgt] grouped by patch, seg_ptr (Q+1) int32.  Returns R (Q,3,3), t (Q,3),
    mask (K) bool, mask_2 (Q) bool, residuals (K), median (Q)."""
    corr = _dev_f32(corr)
    ptr = seg_ptr.to(corr.device, torch.int32).contiguous()
    src, tgt = corr[:, :3].contiguous(), corr[:, 3:6].contiguous()
    R, t, _, res = ops.segmented_kabsch(src, tgt, ptr, eps=1e-6, variant=0, want_res=True)
    med = ops.segmented_median(res, ptr)                                  # torch.median = lower median
    cnt = (ptr[1:] - ptr[:-1]).long()
    seg = torch.repeat_interleave(torch.arange(cnt.numel(), device=corr.device), cnt)
    mask = res < 2.5 * med[seg]                                           # :115
    kept = torch.zeros(cnt.numel(), dtype=torch.float32, device=corr.device).index_add_(0, seg, mask.float())
    mask_2 = kept / cnt.clamp(min=1).float() >= 0.70                      # :117
    return R, t, mask, mask_2, res, med


def refine_local_rigid_correspondences(corr_neigh_2, refine_type='SVD'):
    """(pruned correspondences, T (4,4) float32 cuda, mask, mask_2) -- src/rgb_guided.py:99-125."""
    if refine_type != 'SVD':
        raise NotImplementedError("refine_type 'RANSAC' is Open3D ransac_registration and stays in the reference")
    corr = _dev_f32(corr_neigh_2)
    ptr = torch.tensor([0, corr.shape[0]], dtype=torch.int32, device=corr.device)
    R, t, mask, mask_2, _, _ = refine_local_rigid_correspondences_batched(corr, ptr)
    T = torch.eye(4, device=corr.device)
    T[:3, :3] = R[0]
    T[:3, 3] = t[0]
    return corr[mask], T, mask, mask_2[0]
