"""Host mirror of utils/o3d_tools.py::icp_registration (row E1): Open3D registration_icp point-to-point
(TransformationEstimationPointToPoint(False), ICPConvergenceCriteria(1e-6, 1e-6, 30)) as one persistent
CUDA kernel per batch of patch pairs (f4l_patch_icp), fp64 like Open3D.  utils/o3d_tools.py:12-71."""
import numpy as np
import torch

from . import ops
from .functions import _dev_f32


def _points(pcd, device=None):
    """Accepts an Open3D PointCloud (anything with .points), a numpy array or a tensor."""
    if hasattr(pcd, "points"):
        pcd = np.asarray(pcd.points)
    return _dev_f32(pcd, device)


def icp_registration(src_pcd, tgt_pcd, initial_transform, threshold=0.1, icp_type='point2point'):
    """Returns the reference's dict: fitness, inlier_rmse, correspondence_set (k,2) int, est_transform (4,4)
    float64 numpy, src_corr_pts, tgt_corr_pts (numpy (k,3) instead of Open3D clouds).
    The unneeded normal estimation of the reference (:29-30) is not performed."""
    if icp_type != 'point2point':
        raise ValueError('ICP type not supported') if icp_type not in ('point2plane', 'generalized_icp') else \
            NotImplementedError("only the 'point2point' estimator is on the hot path (base.py:3358, rgb_guided.py:1019)")
    s = _points(src_pcd)
    t = _points(tgt_pcd, s.device)
    T0 = torch.as_tensor(np.asarray(initial_transform.detach().cpu() if torch.is_tensor(initial_transform)
                                    else initial_transform), dtype=torch.float64).reshape(1, 4, 4).to(s.device)
    sp = torch.tensor([0, s.shape[0]], dtype=torch.int32, device=s.device)
    tp = torch.tensor([0, t.shape[0]], dtype=torch.int32, device=s.device)
    T, fit, rmse, iters, corr = ops.patch_icp(s, t, sp, tp, T0=T0, max_corr_dist=threshold, want_corr=True)
    corr = corr.cpu().numpy()
    src_i = np.nonzero(corr >= 0)[0]
    cset = np.stack([src_i, corr[src_i]], axis=1).astype(np.int32)
    s_np, t_np = s.cpu().numpy(), t.cpu().numpy()
    return {
        "fitness": float(fit.item()),
        "inlier_rmse": float(rmse.item()),
        "correspondence_set": cset,
        "est_transform": T[0].cpu().numpy(),
        "src_corr_pts": s_np[cset[:, 0]],
        "tgt_corr_pts": t_np[cset[:, 1]],
    }
