"""CPU: pins the oracle (oracle/*.py, fp64 restatements) to the golden vectors produced by the
UNMODIFIED reference functions (oracle/make_golden.py).  Tolerances are the fp32 rounding of the
reference itself (it computes in fp32, the oracle in fp64)."""
import os

import numpy as np
import pytest

from oracle import desc_nn, knn, rigid


def _cases(z, suffix="_src"):
    return sorted(k[:-len(suffix)] for k in z.files if k.endswith(suffix))


def _act_err(Ra, ta, Rb, tb, pts):
    """max_p || (Ra p + ta) - (Rb p + tb) ||  -- transforms compared by their action."""
    pa = pts.astype(np.float64) @ np.asarray(Ra, np.float64).T + np.asarray(ta, np.float64)
    pb = pts.astype(np.float64) @ np.asarray(Rb, np.float64).T + np.asarray(tb, np.float64)
    return np.linalg.norm(pa - pb, axis=1).max()


def test_procrustes_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "rigid_procrustes.npz"))
    for key in _cases(z):
        s, t, w = z[key + "_src"], z[key + "_tgt"], z[key + "_w"]
        w = None if w.size == 0 else w
        R, tr = rigid.weighted_procrustes(s, t, w, 0.0, eps=1e-6)
        # fp32 reference: ~1e-6 relative on R entries, a few f32 ulps of |p| (<= 60 m) in action
        assert _act_err(R, tr, z[key + "_R"], z[key + "_t"], s) < 5e-5, key
        assert abs(np.linalg.det(R) - 1) < 1e-9, key


def test_kabsch_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "rigid_kabsch.npz"))
    for key in _cases(z):
        s, t, w = z[key + "_src"], z[key + "_tgt"], z[key + "_w"]
        w = None if w.size == 0 else w
        R, tr, res, flag = rigid.kabsch(s, t, w)
        assert not flag
        assert _act_err(R, tr, z[key + "_R"], z[key + "_t"], s) < 5e-5, key
        np.testing.assert_allclose(res, z[key + "_res"], atol=5e-5)
        np.testing.assert_allclose(rigid.transform_point_cloud(s, R, tr), z[key + "_x1t"], atol=5e-5)


def test_filter_tail_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "f2s3_filter.npz"))
    assert "All keys matched" in str(z["load_report"][0]) or "missing_keys=[]" in str(z["load_report"][0])
    n_robust = 0
    for key in _cases(z, "_corr"):
        corr, scores = z[key + "_corr"], z[key + "_scores"]
        coeff = 2.5 if key.endswith("c25") else 1.0
        o = rigid.filter_input_tail(corr[:, :3], corr[:, 3:], scores, coeff)
        assert o["robust_estimate"] == bool(z[key + "_robust"][0]), key
        n_robust += o["robust_estimate"]
        assert _act_err(o["rot_est"], o["trans_est"], z[key + "_R"], z[key + "_t"], corr[:, :3]) < 1e-4, key
    assert n_robust > 0


def test_knn_matches_sklearn_calls(golden_dir):
    z = np.load(os.path.join(golden_dir, "knn_sklearn.npz"))
    a, b = z["a"], z["b"]
    idx, d2 = knn.knn_exact(a, a, 2)
    ties = knn.tie_rows(a, a, 2)
    same = (idx == z["self_i_a"]).all(1)
    assert (same | ties).all()
    np.testing.assert_allclose(np.sqrt(d2), z["self_d_a"], rtol=1e-5, atol=1e-7)
    assert abs(knn.median_resolution(a, b) - z["median_resolution"][0]) < 1e-7
    np.testing.assert_allclose(knn.compute_c2c(a, b), z["c2c"], rtol=1e-9, atol=1e-12)
    i1, _ = knn.knn_exact(a, b, 1)
    t1 = knn.tie_rows(a, b, 1)
    assert ((i1[:, 0] == z["c2c_idx"][:, 0]) | t1).all()


def test_desc_nn_matches_cdist_min(golden_dir):
    z = np.load(os.path.join(golden_dir, "desc_cdist.npz"))
    for D in (32, 64):
        a, b = z["D%d_a" % D], z["D%d_b" % D]
        idx, d2, d2b = desc_nn.desc_nn(a, b, return_second=True)
        tie = (d2b - d2) <= desc_nn.EPS_DESC_ABS
        ref = z["D%d_labels" % D]
        # the reference's fp32 GEMM-formulation cdist can flip near-ties: those rows are exempt
        assert ((idx == ref) | tie).all()
        assert (idx != ref).sum() <= tie.sum()
        ok = idx == ref
        np.testing.assert_allclose(np.sqrt(d2[ok]), z["D%d_dist" % D][ok], atol=2e-3)
    m, j = desc_nn.coarse_matching_3d(z["coarse_cs"], z["coarse_fs"], z["coarse_ct"], z["coarse_ft"],
                                      float(z["coarse_max_mag"][0]), "nn_mutual")
    mask_ref = z["coarse_mutual"] & z["coarse_in_mag"]
    np.testing.assert_array_equal(m, np.nonzero(mask_ref)[0])
    np.testing.assert_array_equal(j, z["coarse_j"][mask_ref])


def test_rigidity_matches_cdist_expression(golden_dir):
    z = np.load(os.path.join(golden_dir, "rigidity_cdist.npz"))
    for ci in range(5):
        r, m = rigid.rigidity_check(z["r%d_src" % ci], z["r%d_tgt" % ci], 0.5)
        # torch.cdist switches to the fp32 GEMM formulation above 25 rows: noise ~1e-3 at |p|~40 m
        assert abs(m - z["r%d_mean" % ci][0]) < 5e-3, ci
        assert abs(r - z["r%d_ratio" % ci][0]) < 5e-3, ci


def test_dips_oracle_matches_reference_golden(golden_dir):
    """oracle/dips.py against the reference's own Preprocess_Dataset output (make_golden.make_dips)."""
    from oracle import dips as odips
    z = np.load(os.path.join(golden_dir, "dips_patches.npz"))
    ref, radius = z["ref"], float(z["radius"][0])
    out, cnt, lrf = odips.patches(ref[z["pick"]], ref, radius, z["inds"])
    np.testing.assert_array_equal(cnt, z["count"])
    np.testing.assert_allclose(out, z["patches"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(lrf, z["lrf"], rtol=0, atol=1e-12)


def test_lifting_oracle_matches_reference_golden(golden_dir):
    """oracle/lifting.py against the reference's own map_corr_2d_to_3d / _tgt2src (make_golden.make_lifting)."""
    from oracle import lifting as olift
    z = np.load(os.path.join(golden_dir, "map_corr_2d.npz"))
    for rev, tag in ((False, "fwd"), (True, "rev")):
        i, m, rows = olift.map_corr_2d_to_3d(z["corres_2d"], z["src_pixel"], z["tgt_pixel"], float(z["thres"][0]), reverse=rev)
        np.testing.assert_array_equal(i, z["idx_" + tag])
        np.testing.assert_array_equal(m, z["mask_" + tag])
        np.testing.assert_array_equal(rows, z["rows_" + tag])
        assert 0 < m.sum() < m.size


def test_voxel_oracle_equals_sequential_hash_map_loop():
    """oracle/voxel.py (vectorised) against a literal transcription of Open3D's VoxelDownSample loop: a hash map from
    the integer voxel index to (sum, count), filled point by point; bit-equal means, same membership."""
    import math
    from oracle import voxel as ovox
    rng = np.random.default_rng(4)
    pts = np.vstack([rng.uniform(-3, 7, (4000, 3)), np.round(rng.uniform(-3, 7, (500, 3)) / 0.25) * 0.25])
    for voxel in (0.25, 0.4, 3.0):
        cent, inv = ovox.voxel_down_sample(pts, voxel)
        vmb = pts.min(0) - voxel * 0.5
        acc = {}
        for p in pts:
            key = tuple(int(math.floor((p[c] - vmb[c]) / voxel)) for c in range(3))
            s_, n_ = acc.get(key, (np.zeros(3), 0))
            acc[key] = (s_ + p, n_ + 1)
        keys = sorted(acc)
        assert len(keys) == cent.shape[0]
        want = np.array([acc[k][0] / acc[k][1] for k in keys])
        np.testing.assert_array_equal(cent, want)
        row_of = {k: i for i, k in enumerate(keys)}
        got_rows = [row_of[tuple(int(math.floor((p[c] - vmb[c]) / voxel)) for c in range(3))] for p in pts]
        np.testing.assert_array_equal(inv, np.array(got_rows))


def _lists_from_labels(lab, min_pts):
    u, c = np.unique(lab, return_counts=True)
    keep = u[c > min_pts]
    return keep, [np.nonzero(lab == k)[0] for k in keep]


def test_coarse_vote_and_mutual_match_reference_method_body(golden_dir):
    """B4 + B3 + pair order: oracle/desc_nn.py against `Coarse2Fine_Base.coarse_matching_with_different_types` itself
    (make_golden.make_coarse ran the unmodified method on a stand-in self).  Source patches with a tied top vote are
    exempt (torch.argsort leaves their order unspecified): 3 forced + whatever the draw produced."""
    z = np.load(os.path.join(golden_dir, "coarse_method.npz"))
    lab_s, lab_t, corr2d = z["lab_s"], z["lab_t"], z["corr2d"]
    idx_spt_src, spt_s = _lists_from_labels(lab_s, int(z["min_pts"][0]))
    idx_spt_tgt, spt_t = _lists_from_labels(lab_t, int(z["min_pts"][0]))
    m2, j2, tie2 = desc_nn.coarse_matching_2d_vote(corr2d, lab_t, spt_s, idx_spt_tgt)
    m3, j3 = desc_nn.coarse_matching_3d(z["cs"], z["fs"], z["ct"], z["ft"], float(z["max_mag"][0]), "nn_mutual")
    tie = z["tie"]
    assert tie.sum() >= 3 and tie2.sum() >= 1
    np.testing.assert_array_equal(tie[m2], tie2)
    n2 = int(z["fusion_spt_length"][0])
    assert z["fusion_spt_length"].tolist() == [n2, int(z["only_3d_spt_length"][0])]
    # 3D pairs: exact, and appended AFTER the 2D pairs in fusion mode
    np.testing.assert_array_equal(z["only_3d_m"], m3)
    np.testing.assert_array_equal(z["only_3d_j"], j3)
    np.testing.assert_array_equal(z["fusion_m"][n2:], m3)
    np.testing.assert_array_equal(z["fusion_j"][n2:], j3)
    np.testing.assert_array_equal(z["fusion_m"][:n2], z["only_2d_m"])
    # 2D pairs: identical on every source patch without a tie
    ref = dict(zip(z["only_2d_m"].tolist(), z["only_2d_j"].tolist()))
    mine = dict(zip(m2.tolist(), j2.tolist()))
    for m in range(len(spt_s)):
        if not tie[m]:
            assert ref.get(m, -1) == mine.get(m, -1), m
    order_ref = [m for m in z["only_2d_m"].tolist() if not tie[m]]
    assert order_ref == [m for m in m2.tolist() if not tie[m]]            # ascending source patch order


def test_merge_matches_reference_function(golden_dir):
    """M1: oracle merge against the reference's own function run with an exact index (make_golden.make_merge)."""
    z = np.load(os.path.join(golden_dir, "merge_levels.npz"))
    levels = [z["level0"], z["level1"], z["level2"]]
    merged, masks = desc_nn.merge_by_priority(levels)
    np.testing.assert_array_equal(merged, z["merged"])
    assert 0 < (~masks[2]).sum() < masks[2].size
    merged0, _ = desc_nn.merge_by_priority([levels[0][:0], levels[1], levels[2]])
    np.testing.assert_array_equal(merged0, z["merged_empty0"])


def test_icp_umeyama_step_matches_opencv():
    """Partial pin of the (otherwise unpinned) ICP oracle: its rigid-update step, Eigen::umeyama without scaling as restated
    in oracle/icp.py, against an independent compiled implementation of the same published algorithm -- OpenCV's
    cv2.estimateAffine3D(src, dst, force_rotation=True) (calib3d, Umeyama 1991).  The rotation does not depend on the
    scale estimate, so the rotations must agree to rounding, including improper optimal maps (a mirrored target: the
    last singular vector is flipped) and rank-deficient covariances (coplanar points)."""
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "estimateAffine3D"):
        pytest.skip("this OpenCV build has no estimateAffine3D")
    from scipy.spatial.transform import Rotation
    from oracle import icp as oicp
    rng = np.random.default_rng(1)
    worst = 0.0
    for case in range(200):
        n = int(rng.integers(3, 200))
        src = rng.random((n, 3)) * rng.uniform(0.5, 20) + rng.uniform(-100, 100, 3)
        R = Rotation.from_rotvec(rng.normal(size=3) * rng.uniform(0, 1.5)).as_matrix()
        t = rng.normal(size=3)
        dst = src @ R.T + t + rng.uniform(0, 0.05) * rng.standard_normal((n, 3))
        if case % 4 == 1:
            dst = dst * np.array([1.0, 1.0, -1.0])
        if case % 4 == 2:
            src[:, 2] = src[0, 2]
            dst = src @ R.T + t
        T = oicp.umeyama_noscale(src, dst)
        try:
            Rt, _ = cv2.estimateAffine3D(src, dst, force_rotation=True)
        except TypeError:
            pytest.skip("cv2.estimateAffine3D without the Umeyama overload")
        assert abs(np.linalg.det(T[:3, :3]) - 1.0) < 1e-9
        worst = max(worst, float(np.abs(T[:3, :3] - Rt[:, :3]).max()))
        # the translation of the no-scale variant follows from the rotation and the means
        np.testing.assert_allclose(T[:3, 3], dst.mean(0) - T[:3, :3] @ src.mean(0), atol=1e-9)
    assert worst < 1e-9, worst
