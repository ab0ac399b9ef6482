"""GPU parity: K-e per-patch point-to-point ICP and the segmented 1-NN (row A4), through the C ABI,
against the fp64 restatement of Open3D's registration_icp (oracle/icp.py -- parity unpinned: Open3D
is not installable here) and the cKDTree restatement of refine_dvfs_with_threshold."""
import numpy as np
import pytest
import torch

from oracle import icp as oicp
from oracle import knn as oknn
from oracle import rigid
from tests.test_oracle_golden import _act_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _tile_patches(n_pts=60_000, seed=21, patch_pts=256):
    from fusion4landslide_b200 import synth
    d = synth.make_tile(n_pts, seed=seed, patch_pts=patch_pts)
    lab_s, ptr_s, idx_s = synth.patches_from_labels(d["label_src"])
    lab_t, ptr_t, idx_t = synth.patches_from_labels(d["label_tgt"])
    m, j = synth.pair_patches(lab_s, lab_t)
    return d, (ptr_s, idx_s), (ptr_t, idx_t), m, j


def _matched(d, ps, pt):
    c = d["corr3d"].numpy()[ps]
    c = c[np.isin(c[:, 1], pt)]
    return d["src"].numpy()[c[:, 0]], d["tgt"].numpy()[c[:, 1]]


@pytest.mark.parametrize("init", ["identity", "procrustes"])
def test_patch_icp_vs_oracle(cuda, init):
    from fusion4landslide_b200 import ops
    d, (ptr_s, idx_s), (ptr_t, idx_t), m, j = _tile_patches()
    A_list, B_list, T0_list = [], [], []
    for a, b in zip(m.tolist()[:150], j.tolist()[:150]):
        ps = idx_s[ptr_s[a]:ptr_s[a + 1]].numpy()
        pt = idx_t[ptr_t[b]:ptr_t[b + 1]].numpy()
        A, B = _matched(d, ps, pt)
        if A.shape[0] < 10:
            continue
        A_list.append(A)
        B_list.append(B)
        T0_list.append(np.eye(4) if init == "identity" else rigid.procrustes_transform(A, B, eps=1e-6))
    # add an empty patch and one larger than the shared-memory staging capacity
    big = np.concatenate(A_list[:30]), np.concatenate(B_list[:30])
    A_list += [np.zeros((0, 3), np.float32), big[0]]
    B_list += [np.zeros((0, 3), np.float32), big[1]]
    T0_list += [np.eye(4), np.eye(4)]
    Q = len(A_list)
    sptr = np.zeros(Q + 1, np.int32)
    sptr[1:] = np.cumsum([len(a) for a in A_list])
    src = torch.from_numpy(np.concatenate(A_list)).to(cuda)
    tgt = torch.from_numpy(np.concatenate(B_list)).to(cuda)
    dptr = torch.from_numpy(sptr).to(cuda)
    T0 = torch.from_numpy(np.stack(T0_list)).to(cuda)
    T, fit, rmse, iters, corr = ops.patch_icp(src, tgt, dptr, dptr, T0=T0, max_corr_dist=0.1, want_corr=True)
    torch.cuda.synchronize()
    T, fit, rmse, iters, corr = (x.cpu().numpy() for x in (T, fit, rmse, iters, corr))
    n_flip = n_degenerate = 0
    for q in range(Q):
        o = oicp.icp_point_to_point(A_list[q], B_list[q], T0_list[q], 0.1)
        if 0 < o["min_ncorr"] < 4 or (o["min_ncorr"] > 0 and o["min_sv_ratio"] < 1e-6):
            n_degenerate += 1   # rigid fit from < 4 pairs / collinear pairs is not unique: exempt
            continue
        same_path = iters[q] == o["iters"] and abs(fit[q] - o["fitness"]) < 1e-12
        if not same_path:
            n_flip += 1   # a nearest-neighbour / inlier decision within fp64 rounding of a tie
            continue
        assert abs(rmse[q] - o["inlier_rmse"]) < 1e-9, q
        if len(A_list[q]):
            assert _act_err(T[q, :3, :3], T[q, :3, 3], o["transformation"][:3, :3], o["transformation"][:3, 3],
                            A_list[q]) < TOL, q
        c = corr[sptr[q]:sptr[q + 1]]
        oc = -np.ones(len(A_list[q]), np.int64)
        oc[o["correspondence_set"][:, 0]] = o["correspondence_set"][:, 1]
        # duplicate target points (several src matched to one tgt) are exact ties: compare coordinates
        ok = (c >= 0) == (oc >= 0)
        assert ok.all(), q
        sel = c >= 0
        np.testing.assert_array_equal(B_list[q][c[sel]], B_list[q][oc[sel]])
    print("icp: %d patches, %d degenerate (exempt), %d near-tie path flips" % (Q, n_degenerate, n_flip))
    assert n_flip <= 1, "%d of %d patches took a different ICP path" % (n_flip, Q)
    assert n_degenerate <= Q // 8
    assert iters.max() <= 30 and (iters[:-2] >= 1).all()


def test_patch_icp_skip_and_limits(cuda):
    from fusion4landslide_b200 import ops
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 2, size=(40, 3)).astype(np.float32)
    b = (a + 0.03).astype(np.float32)
    src = torch.from_numpy(np.concatenate([a, a])).to(cuda)
    tgt = torch.from_numpy(np.concatenate([b, b])).to(cuda)
    ptr = torch.tensor([0, 40, 80], dtype=torch.int32, device=cuda)
    skip = torch.tensor([1, 0], dtype=torch.uint8, device=cuda)
    T, fit, rmse, iters = ops.patch_icp(src, tgt, ptr, ptr, max_corr_dist=0.5, seg_skip=skip)
    assert iters.tolist()[0] == 0 and torch.equal(T[0], torch.eye(4, dtype=torch.float64, device=cuda))
    o = oicp.icp_point_to_point(a, b, None, 0.5)
    assert iters.tolist()[1] == o["iters"]
    # max_iter = 0 -> only the initial match
    T, fit, rmse, iters = ops.patch_icp(src, tgt, ptr, ptr, max_corr_dist=0.5, max_iter=0)
    o0 = oicp.icp_point_to_point(a, b, None, 0.5, max_iter=0)
    assert iters.tolist() == [0, 0] and abs(fit[0].item() - o0["fitness"]) < 1e-12
    with pytest.raises(RuntimeError):
        ops.patch_icp(src, tgt, ptr, ptr, max_corr_dist=0.0)


def _crafted_fragile_cases():
    """(source, target, max_dist, expected flag bit) -- exact in f32, so the decisions really sit ON the boundary."""
    rng = np.random.default_rng(11)
    base = (rng.integers(0, 64, size=(40, 3)) * 0.25).astype(np.float32)          # a rigid, well separated cloud
    cases = []
    # (a) one source point exactly midway between two DISTINCT targets (identity start, exact match elsewhere)
    a = np.concatenate([base, [[100.0, 0.0, 0.0]]]).astype(np.float32)
    b = np.concatenate([base, [[99.5, 0.0, 0.0], [100.5, 0.0, 0.0]]]).astype(np.float32)
    cases.append((a, b, 1.0, 1))
    # (b) a nearest distance exactly equal to max_correspondence_distance (strict `<`: rejected, by a hair)
    a = np.concatenate([base, [[100.0, 0.0, 0.0]]]).astype(np.float32)
    b = np.concatenate([base, [[100.5, 0.0, 0.0]]]).astype(np.float32)
    cases.append((a, b, 0.5, 2))
    # (c) duplicates of the winner are NOT a tie (same coordinates: the transform cannot depend on which one is taken)
    a = base.copy()
    b = np.concatenate([base, base[:7]]).astype(np.float32)
    cases.append((a, b, 1.0, 0))
    return cases


@pytest.mark.parametrize("pad", [0, 300])     # 300 extra far-away targets push the pair onto the CTA kernel (> 224 points)
def test_patch_icp_fragile_flags_crafted(cuda, pad):
    from fusion4landslide_b200 import ops
    A_list, B_list, want = [], [], []
    far = (np.arange(pad, dtype=np.float32)[:, None] * np.array([[0.0, 3.0, 0.0]], np.float32) + np.array([[0, 1000, 0]], np.float32))
    for a, b, md, bit in _crafted_fragile_cases():
        A_list.append(a)
        B_list.append(np.concatenate([b, far]).astype(np.float32) if pad else b)
        want.append((md, bit))
    for (md, bit), a, b in zip(want, A_list, B_list):
        src, tgt = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
        sp = torch.tensor([0, len(a)], dtype=torch.int32, device=cuda)
        tp = torch.tensor([0, len(b)], dtype=torch.int32, device=cuda)
        out = ops.patch_icp(src, tgt, sp, tp, max_corr_dist=md, want_fragile=True)
        flag = int(out[-1].item())
        if bit:
            assert flag & bit, (md, bit, flag)
        else:
            assert flag & 3 == 0, (md, bit, flag)


def test_patch_icp_path_flips_only_on_fragile_pairs(cuda):
    """The tie[Q] contract: a pair whose ICP path differs from the fp64 restatement must carry the flag (the converse does
    not hold -- most fragile decisions still go the same way)."""
    from fusion4landslide_b200 import ops
    d, (ptr_s, idx_s), (ptr_t, idx_t), m, j = _tile_patches(n_pts=80_000, seed=5)
    A_list, B_list = [], []
    for a, b in zip(m.tolist()[:250], j.tolist()[:250]):
        ps = idx_s[ptr_s[a]:ptr_s[a + 1]].numpy()
        pt = idx_t[ptr_t[b]:ptr_t[b + 1]].numpy()
        A, B = _matched(d, ps, pt)
        if A.shape[0] >= 10:
            A_list.append(A)
            B_list.append(B)
    Q = len(A_list)
    sptr = np.zeros(Q + 1, np.int32)
    sptr[1:] = np.cumsum([len(a) for a in A_list])
    src = torch.from_numpy(np.concatenate(A_list)).to(cuda)
    tgt = torch.from_numpy(np.concatenate(B_list)).to(cuda)
    dptr = torch.from_numpy(sptr).to(cuda)
    T, fit, rmse, iters, fragile = ops.patch_icp(src, tgt, dptr, dptr, max_corr_dist=0.1, want_fragile=True)
    iters, fit, fragile = iters.cpu().numpy(), fit.cpu().numpy(), fragile.cpu().numpy()
    n_flip = 0
    for q in range(Q):
        o = oicp.icp_point_to_point(A_list[q], B_list[q], None, 0.1)
        if 0 < o["min_ncorr"] < 4 or (o["min_ncorr"] > 0 and o["min_sv_ratio"] < 1e-6):
            continue
        if not (iters[q] == o["iters"] and abs(fit[q] - o["fitness"]) < 1e-12):
            n_flip += 1
            assert fragile[q] != 0, "pair %d left the oracle's path without a fragility flag" % q
    print("icp fragility: %d pairs, %d flagged, %d path flips" % (Q, int((fragile != 0).sum()), n_flip))
    assert (fragile != 0).mean() < 0.2


def test_segmented_nn_vs_oracle(cuda):
    from fusion4landslide_b200 import ops
    d, (ptr_s, idx_s), (ptr_t, idx_t), m, j = _tile_patches(n_pts=40_000, seed=22)
    m, j = m[:120], j[:120]
    Q = m.numel()
    # CSR over pairs
    def sub(ptr, idx, sel):
        cnt = (ptr[1:] - ptr[:-1])[sel]
        p = torch.zeros(sel.numel() + 1, dtype=torch.int32)
        p[1:] = torch.cumsum(cnt, 0)
        items = torch.cat([idx[ptr[a]:ptr[a + 1]] for a in sel.tolist()])
        return p, items
    sp, si = sub(ptr_s, idx_s, m)
    tp, ti = sub(ptr_t, idx_t, j)
    rng = np.random.default_rng(4)
    T = np.tile(np.eye(4, dtype=np.float32), (Q, 1, 1))
    T[:, :3, 3] = rng.normal(size=(Q, 3)) * 0.05
    thr = rng.uniform(0.03, 0.2, size=Q).astype(np.float32)
    nn, d2 = ops.segmented_nn(d["src"].to(cuda), d["tgt"].to(cuda), sp.to(cuda), tp.to(cuda), qidx=si.to(cuda),
                              ridx=ti.to(cuda), T=torch.from_numpy(T).to(cuda), thr=torch.from_numpy(thr).to(cuda))
    torch.cuda.synchronize()
    nn, d2 = nn.cpu().numpy(), d2.cpu().numpy()
    src, tgt = d["src"].numpy(), d["tgt"].numpy()
    bad = 0
    for q in range(Q):
        S = src[si[sp[q]:sp[q + 1]].numpy()]
        Tg = tgt[ti[tp[q]:tp[q + 1]].numpy()]
        moved = (S.astype(np.float64) @ T[q, :3, :3].astype(np.float64).T + T[q, :3, 3].astype(np.float64)).astype(np.float32)
        rows, keep, onn = oknn.refine_dvfs_with_threshold(S, moved, Tg, float(thr[q]))
        got = nn[sp[q]:sp[q + 1]]
        assert ((got >= 0) == keep).all(), q
        bad += int((got[keep] != onn[keep]).sum())
    assert bad == 0
