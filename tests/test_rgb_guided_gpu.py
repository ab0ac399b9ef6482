"""GPU parity of the RGB-guided per-patch loop, `Image_DVFs.local_rigid_refinement` (src/rgb_guided.py:981-1062), run for
all segment patches at once (rgb_guided.local_rigid_refinement_batched and the mixin that writes the reference's
`data_output` fields), against the patch-by-patch oracle restatement (oracle/paths.rgb_local_rigid_refinement)."""
import types

import numpy as np
import pytest
import torch

from oracle import paths as opaths

pytestmark = pytest.mark.gpu


def _case(n=24_000, seed=4):
    from fusion4landslide_b200 import synth
    d = synth.make_tile(n, seed=seed, patch_pts=180)              # 1:1 counterparts: every source point has a lifted match
    src, tgt = d["src"].numpy(), d["tgt"].numpy()
    rng = np.random.default_rng(seed)
    valid = np.sort(rng.choice(n, int(0.8 * n), replace=False))   # idx_valid_src: source points with a 2D-lifted match
    t_of = d["gt_tgt_of_src"].numpy()[valid]
    wrong = rng.random(valid.size) < 0.05
    t_of[wrong] = rng.integers(0, n, int(wrong.sum()))
    corr = np.hstack([src[valid], tgt[t_of]]).astype(np.float32)
    lab = d["label_src"].numpy()
    u, c = np.unique(lab, return_counts=True)
    patches = [np.nonzero(lab == k)[0] for k in u[c > 10]]          # ids of ALL points of the patch (valid or not)
    patches.append(np.array([n + 5, n + 6], np.int64))            # a patch none of whose points has a match
    return corr, valid, t_of, patches


def test_local_rigid_refinement_batched_vs_oracle(cuda):
    from fusion4landslide_b200 import rgb_guided
    corr, valid, _, patches = _case()
    r = rgb_guided.local_rigid_refinement_batched(torch.from_numpy(corr).to(cuda), torch.from_numpy(valid).to(cuda),
                                                  [torch.from_numpy(p) for p in patches], icp_thres=0.1)
    torch.cuda.synchronize()
    keep_o, rows_o, per = opaths.rgb_local_rigid_refinement(corr, valid, patches, icp_thres=0.1)
    np.testing.assert_array_equal(r["mask_valid_local"].cpu().numpy(), keep_o)
    rows = r["corres_3d_refine_apply_icp"].cpu().numpy()
    assert rows.shape == rows_o.shape and rows.shape[0] > 15_000
    np.testing.assert_array_equal(rows[:, :3], rows_o[:, :3])
    ptr = r["seg_ptr"].cpu().numpy()
    iters = r["iters"].cpu().numpy()
    assert ptr[-1] == rows.shape[0] and ptr[-1] - ptr[-2] == 0                          # the empty patch contributes nothing
    flips = degenerate = checked = 0
    for q, p in enumerate(per):
        if p is None:
            continue
        a, b = ptr[q], ptr[q + 1]
        assert b - a == p["n"]
        if round(min(p["fitness"], p["fitness0"]) * p["n"]) < 3:
            # < 3 inlier correspondences at the start or at the end: the Umeyama problem of that iteration is rank
            # deficient, every SVD returns another valid minimiser and the trajectories part (patches whose Procrustes
            # start was ruined by gross mismatches)
            degenerate += 1
            continue
        if iters[q] != p["iters"]:
            flips += 1
            continue
        checked += 1
        assert np.abs(rows[a:b] - rows_o[a:b]).max() < 1e-5 + 2 * np.spacing(np.float32(np.abs(rows_o).max())), q
    assert flips <= max(1, len(per) // 100), flips
    assert degenerate <= max(2, len(per) // 10) and checked > 0.88 * (len(per) - 1), (degenerate, checked, len(per))


def test_image_dvfs_mixin_writes_the_reference_fields(cuda):
    from fusion4landslide_b200 import rgb_guided
    from fusion4landslide_b200.entry_c2f import edict
    corr, valid, t_of, patches = _case(12_000, 9)

    class Base:                                                     # what the mixin needs of Image_DVFs
        def __init__(self):
            self.verbose, self.logging = False, None
            self.method = edict(icp_refine=True, icp_thres=0.1)
            self.data_interim = edict(segment_patches=[torch.from_numpy(p) for p in patches])
            c = torch.from_numpy(corr).to(cuda)
            self.data_output = edict(corres_3d_refine=c, idx_valid_src_refine=torch.from_numpy(valid).to(cuda),
                                     idx_valid_tgt_refine=torch.from_numpy(t_of).to(cuda),
                                     corres_3d_magnitude_refine=torch.linalg.norm(c[:, 3:6] - c[:, :3], dim=1)[:, None])

    obj = rgb_guided.bind(Base)()
    obj.local_rigid_refinement()
    keep_o, rows_o, _ = opaths.rgb_local_rigid_refinement(corr, valid, patches, icp_thres=0.1)
    do = obj.data_output
    np.testing.assert_array_equal(do.idx_valid_src_refine.cpu().numpy(), valid[keep_o])
    np.testing.assert_array_equal(do.idx_valid_tgt_refine.cpu().numpy(), t_of[keep_o])
    np.testing.assert_array_equal(do.corres_3d_refine.cpu().numpy(), corr[keep_o])
    assert do.corres_3d_magnitude_refine.shape == (keep_o.size, 1)
    assert do.corres_3d_refine_apply_icp.shape == rows_o.shape
    assert do.corres_3d_magnitude_refine_apply_icp.shape == (rows_o.shape[0], 1)
