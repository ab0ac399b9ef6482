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


def local_rigid_refinement_batched(corres_3d_refine, idx_valid_src_refine, segment_patches, icp_thres=0.1, icp_refine=True,
                                   icp_max_iter=30):
    """`Image_DVFs.local_rigid_refinement` (src/rgb_guided.py:981-1062) for all segment patches at once.

    corres_3d_refine (n,6) rows [src | tgt]; idx_valid_src_refine (n) the source point id of every row (unique);
    segment_patches: list of 1-D arrays / tensors of source point ids (`data_interim.segment_patches`).
    Per patch, as the reference's loop: rows of the patch (ids absent from idx_valid_src_refine are skipped, patch order
    kept) -> Procrustes + `res < 2.5 * median` mask (refine_local_rigid_correspondences) -> point-to-point ICP between ALL
    source and target points of the patch, initialised with the Procrustes fit (:1004-1019) -> rows [src | T_icp src]
    (:1027-1046).  Returns a dict: mask_valid_local (rows kept, patch order -- what :1050-1055 index with),
    corres_3d_refine_apply_icp (rows of every non-empty patch back to back), T_initial (Q,4,4), T_icp (Q,4,4) f32,
    mask_robust (Q) bool, seg_ptr (Q+1) i32, rows (K) i64 (row of corres_3d_refine behind every patch item)."""
    corr_all = _dev_f32(corres_3d_refine)
    dev = corr_all.device
    ids = torch.as_tensor(idx_valid_src_refine).to(dev, torch.int64).reshape(-1)
    if len(segment_patches):
        values = torch.cat([torch.as_tensor(p).reshape(-1).to(torch.int64) for p in segment_patches]).to(dev)
        lens = torch.tensor([int(torch.as_tensor(p).numel()) for p in segment_patches], dtype=torch.int64, device=dev)
    else:
        values = torch.zeros(0, dtype=torch.int64, device=dev)
        lens = torch.zeros(0, dtype=torch.int64, device=dev)
    Q = int(lens.numel())
    size = int(max(int(ids.max().item()) if ids.numel() else -1, int(values.max().item()) if values.numel() else -1)) + 1
    lookup = torch.full((max(size, 1),), -1, dtype=torch.int64, device=dev)
    lookup[ids] = torch.arange(ids.numel(), device=dev)
    rows_all = lookup[values]                                                          # :990 torch.where(... == value)
    present = rows_all >= 0
    seg_all = torch.repeat_interleave(torch.arange(Q, device=dev), lens, output_size=int(values.numel()))
    cnt = torch.zeros(Q, dtype=torch.int64, device=dev).index_add_(0, seg_all[present], torch.ones_like(seg_all[present]))
    ptr = torch.zeros(Q + 1, dtype=torch.int32, device=dev)
    ptr[1:] = torch.cumsum(cnt, 0).to(torch.int32)
    rows = rows_all[present]
    corr = corr_all[rows].contiguous()
    out = {"seg_ptr": ptr, "rows": rows}
    if corr.shape[0] == 0:
        out.update(mask_valid_local=rows, corres_3d_refine_apply_icp=corr.reshape(0, 6),
                   T_initial=torch.eye(4, device=dev).repeat(Q, 1, 1), T_icp=torch.eye(4, device=dev).repeat(Q, 1, 1),
                   mask_robust=torch.zeros(Q, dtype=torch.bool, device=dev))
        return out
    R, t, mask, mask_2, _, _ = refine_local_rigid_correspondences_batched(corr, ptr)
    T0 = torch.eye(4, dtype=torch.float64, device=dev).repeat(Q, 1, 1)
    T0[:, :3, :3] = R.double()
    T0[:, :3, 3] = t.double()
    out["T_initial"] = T0.float()
    out["mask_robust"] = mask_2
    out["mask_valid_local"] = rows[mask]                                               # :1001
    if icp_refine:
        src, tgt = corr[:, :3].contiguous(), corr[:, 3:6].contiguous()
        empty = (cnt == 0).to(torch.uint8)                                             # :1003 temp_corr.shape[0] > 0
        T64, fit, rmse, iters = ops.patch_icp(src, tgt, ptr, ptr, T0=T0.reshape(Q, 16).contiguous(),
                                              max_corr_dist=float(icp_thres), max_iter=int(icp_max_iter), seg_skip=empty)
        T32 = T64.float()                                                              # :1024 dtype=torch.float32
        dvf, _ = ops.apply_transforms(src, ptr, T32.contiguous(), want_mag=False)      # :1027-1030
        out.update(T_icp=T32, corres_3d_refine_apply_icp=dvf, fitness=fit, rmse=rmse, iters=iters)
    return out


class RGBGuidedMixin:
    """`Image_DVFs.local_rigid_refinement` on the kernels, state in the reference's `data_output` / `data_interim` fields."""

    def local_rigid_refinement(self):
        if getattr(self, "verbose", False) and getattr(self, "logging", None) is not None:
            self.logging.info('Start rigid refinement...')
        do = self.data_output
        r = local_rigid_refinement_batched(do.corres_3d_refine, do.idx_valid_src_refine, self.data_interim.segment_patches,
                                           icp_thres=float(self.method.icp_thres), icp_refine=bool(self.method.icp_refine))
        self.rigid_refinement_result = r
        if len(self.data_interim.segment_patches):                                     # :1049-1062
            keep = r["mask_valid_local"]
            as_dev = lambda x: torch.as_tensor(x).to(keep.device)
            do.idx_valid_src_refine = as_dev(do.idx_valid_src_refine)[keep]
            do.idx_valid_tgt_refine = as_dev(do.idx_valid_tgt_refine)[keep]
            do.corres_3d_refine = as_dev(do.corres_3d_refine)[keep, :]
            do.corres_3d_magnitude_refine = as_dev(do.corres_3d_magnitude_refine)[keep]
            if self.method.icp_refine:
                do.corres_3d_refine_apply_icp = r["corres_3d_refine_apply_icp"]
                do.corres_3d_magnitude_refine_apply_icp = torch.linalg.norm(
                    do.corres_3d_refine_apply_icp[:, 3:6] - do.corres_3d_refine_apply_icp[:, :3], dim=1)[:, None]


def bind(base):
    """class Image_DVFs(RGBGuidedMixin, base): the per-patch refinement loop replaced, everything else inherited."""
    class Image_DVFs(RGBGuidedMixin, base):
        pass
    Image_DVFs.__doc__ = "Drop-in for `from src.rgb_guided import Image_DVFs` (main_rgb_guided.py:13,109-111)."
    return Image_DVFs
