"""TEST INFRASTRUCTURE -- fp64 restatement of Open3D 0.19.0 point-to-point ICP.  PARITY UNPINNED.

The reference calls (utils/o3d_tools.py:12-71, from src/coarse_to_fine_matching_base.py:3358
and src/rgb_guided.py:1019)

    o3d.pipelines.registration.registration_icp(source, target, max_correspondence_distance,
        init, TransformationEstimationPointToPoint(False),
        ICPConvergenceCriteria(relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30))

Open3D (requirements.txt:1 pins open3d==0.19.0) is a third-party dependency that is neither
vendored under the reference tree nor installed in this image, and the reference holds no test
or golden vector for this call -> "parity unpinned".  This file restates the published algorithm
of open3d/pipelines/registration/Registration.cpp (RegistrationICP +
GetRegistrationResultAndCorrespondences) and Eigen::umeyama (with_scaling = false):

  T <- init;  P <- T * source
  res <- match(P)                                  # per point: 1-NN in target, accepted iff
  for i in 0..max_iteration-1:                     #   d2 < max_dist^2 (strict, hybrid search)
      U <- umeyama(P[corr.src], target[corr.tgt])  # identity when corr is empty
      T <- U * T;  P <- U * P
      prev <- res;  res <- match(P)
      if |prev.fitness - res.fitness| < rel_fitness and |prev.rmse - res.rmse| < rel_rmse: break
  fitness = #corr / #source ; inlier_rmse = sqrt(sum d2 / #corr)  (both 0 when no corr)

All arithmetic is fp64, as in Open3D (the reference promotes its f32 tensors, o3d_tools.py:180-257).

Partial pin: the rigid-update step (umeyama_noscale) agrees to 3e-14 with OpenCV's independent implementation of the
same algorithm, cv2.estimateAffine3D(force_rotation=True), on 200 random cases including mirrored and coplanar inputs
(tests/test_oracle_golden.py::test_icp_umeyama_step_matches_opencv).  The loop around it (match, strict d2 < max^2,
both-delta convergence test) stays a restatement without a third-party check.
"""
import numpy as np
from scipy.spatial import cKDTree


def umeyama_noscale(src, dst):
    """Eigen::umeyama(src, dst, false) for (n,3) arrays.  Returns 4x4."""
    n = src.shape[0]
    ms = src.mean(0)
    md = dst.mean(0)
    sigma = (dst - md).T @ (src - ms) / n
    U, sv, Vt = np.linalg.svd(sigma)
    umeyama_noscale.last_sv = sv
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1.0
    R = U @ np.diag(S) @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = md - R @ ms
    return T


def _match(P, tree, tgt_n, max_dist):
    d, j = tree.query(P, k=1)
    d2 = d * d
    ok = d2 < max_dist * max_dist
    ncorr = int(ok.sum())
    if ncorr == 0:
        return ok, j, 0.0, 0.0
    return ok, j, ncorr / P.shape[0], float(np.sqrt(d2[ok].sum() / ncorr))


def icp_point_to_point(source, target, init=None, max_dist=0.1, max_iter=30,
                       rel_fitness=1e-6, rel_rmse=1e-6):
    """Returns dict(transformation 4x4, fitness, inlier_rmse, correspondence_set (c,2), iters)."""
    src = np.asarray(source, dtype=np.float64)
    tgt = np.asarray(target, dtype=np.float64)
    T = np.eye(4) if init is None else np.asarray(init, dtype=np.float64).copy()
    if src.shape[0] == 0 or tgt.shape[0] == 0:
        return dict(transformation=T, fitness=0.0, inlier_rmse=0.0,
                    correspondence_set=np.zeros((0, 2), np.int64), iters=0, min_ncorr=0, min_sv_ratio=0.0)
    tree = cKDTree(tgt)
    P = src @ T[:3, :3].T + T[:3, 3]
    ok, j, fit, rmse = _match(P, tree, tgt.shape[0], max_dist)
    it = 0
    # bookkeeping for the tests: the rigid fit is not unique when fewer than 3 pairs (or a
    # rank-deficient covariance) enter it -- those patches are exempt from path comparison
    min_ncorr, min_sv_ratio = int(ok.sum()), 1.0
    for it in range(1, max_iter + 1):
        min_ncorr = min(min_ncorr, int(ok.sum()))
        if ok.any():
            U = umeyama_noscale(P[ok], tgt[j[ok]])
            sv = umeyama_noscale.last_sv
            min_sv_ratio = min(min_sv_ratio, sv[1] / sv[0] if sv[0] > 0 else 0.0)
        else:
            U = np.eye(4)
        T = U @ T
        P = P @ U[:3, :3].T + U[:3, 3]
        pfit, prmse = fit, rmse
        ok, j, fit, rmse = _match(P, tree, tgt.shape[0], max_dist)
        if abs(pfit - fit) < rel_fitness and abs(prmse - rmse) < rel_rmse:
            break
    corr = np.stack([np.nonzero(ok)[0], j[ok]], axis=1).astype(np.int64)
    return dict(transformation=T, fitness=fit, inlier_rmse=rmse, correspondence_set=corr, iters=it,
                min_ncorr=min_ncorr, min_sv_ratio=min_sv_ratio)
