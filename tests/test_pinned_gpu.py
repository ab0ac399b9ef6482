"""GPU parity against goldens produced by the reference's OWN method bodies / models (oracle/make_golden.py:
make_coarse, make_merge, make_nets): B4 2D vote + B3 mutual NN + pair order, M1 level merge, and the two learned
models of SURVEY 8(f)-2 with the shipped weights, evaluated for all patches at once."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _T(x, dev, dt=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    return t if dt is None else t.to(dt)


def test_coarse_matching_vs_reference_method_body(cuda, golden_dir):
    """`coarse_matching_with_different_types` (base.py:2925-3157) in fusion mode: 2D-vote pairs, then mutual 3D
    pairs.  Exact on every source patch without a tied vote; tied patches must be FLAGGED by the kernel."""
    from fusion4landslide_b200 import coarse_to_fine as c2f
    from fusion4landslide_b200 import ops
    z = np.load(os.path.join(golden_dir, "coarse_method.npz"))
    min_pts = int(z["min_pts"][0])
    lab_s, lab_t = _T(z["lab_s"], cuda), _T(z["lab_t"], cuda)
    idx_spt_src, ptr_s, idx_s, _ = ops.labels_to_csr(lab_s, min_pts)
    idx_spt_tgt, ptr_t, idx_t, _ = ops.labels_to_csr(lab_t, min_pts)
    m2, j2, tie2 = c2f.coarse_matching_2d(_T(z["corr2d"], cuda), idx_s, ptr_s, lab_t, idx_spt_tgt)
    m3, j3 = c2f.coarse_matching_3d(z["cs"], z["fs"], z["ct"], z["ft"], float(z["max_mag"][0]), "nn_mutual")
    np.testing.assert_array_equal(m3.cpu().numpy(), z["only_3d_m"])
    np.testing.assert_array_equal(j3.cpu().numpy(), z["only_3d_j"])
    tie = z["tie"]
    m2n, j2n = m2.cpu().numpy(), j2.cpu().numpy()
    np.testing.assert_array_equal(tie2.cpu().numpy(), tie[m2n])
    ref = dict(zip(z["only_2d_m"].tolist(), z["only_2d_j"].tolist()))
    mine = dict(zip(m2n.tolist(), j2n.tolist()))
    n_checked = 0
    for m in range(ptr_s.numel() - 1):
        if not tie[m]:
            assert ref.get(m, -1) == mine.get(m, -1), m
            n_checked += m in ref
    assert n_checked > 30 and tie.sum() >= 3
    assert np.all(np.diff(m2n) > 0)                              # ascending source patch order, as the reference's loop
    # fusion order: 2D pairs first (base.py:3139-3146)
    n2 = int(z["fusion_spt_length"][0])
    np.testing.assert_array_equal(z["fusion_m"][n2:], m3.cpu().numpy())


def test_merge_levels_vs_reference_function(cuda, golden_dir):
    from fusion4landslide_b200 import coarse_to_fine as c2f
    z = np.load(os.path.join(golden_dir, "merge_levels.npz"))
    levels = [_T(z["level%d" % k], cuda) for k in range(3)]
    merged = c2f.merge_correspondences_by_priority_with_distance_threshold(levels)
    np.testing.assert_array_equal(merged.cpu().numpy(), z["merged"])
    merged0 = c2f.merge_correspondences_by_priority_with_distance_threshold([levels[0][:0], levels[1], levels[2]])
    np.testing.assert_array_equal(merged0.cpu().numpy(), z["merged_empty0"])


def _load(module, z, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}
    module.load_state_dict(sd)
    return module.eval()


def test_filtering_network_all_supervoxels_at_once(cuda, golden_dir):
    """nets.FilteringNetwork.compute_weights_segments (one pass over every supervoxel, segment kernels for the two
    normalisations) against the reference network run supervoxel by supervoxel with the shipped weights
    (f2s3.py:340-347, outlier_classifier.py:52-63).  fp32 both sides; 24 normalised layers -> 2e-4 on weights in [0,1)."""
    from fusion4landslide_b200 import nets, ops
    z = np.load(os.path.join(golden_dir, "nets_shipped.npz"))
    net = _load(nets.FilteringNetwork(), z, "filter/").to(cuda)
    corr = _T(z["filter_corr"], cuda)                             # float64 rows, unscaled
    ptr = _T(z["filter_ptr"], cuda)
    with torch.no_grad():
        scaled = ops.segment_scale_maxabs(corr, ptr)              # f64 division, then float (f2s3.py:343,346)
        w = net.compute_weights_segments(scaled, ptr, scale=False)
        # the reference's single-supervoxel interface on the same module (plain torch forward)
        p = z["filter_ptr"]
        w_one = torch.cat([net.compute_weights(scaled[p[q]:p[q + 1]].float()[None, None]).reshape(-1) for q in range(p.size - 1)])
    ref = z["filter_scores"]
    assert np.abs(w.cpu().numpy() - ref).max() < 2e-4
    assert np.abs(w_one.cpu().numpy() - ref).max() < 2e-4
    assert 0.05 < (ref > 0).mean() < 1.0


def test_attention_pooling_all_superpoints_at_once(cuda, golden_dir):
    """nets.ClusterFeatureNetWithAttention.aggregate_segments against the reference's per-superpoint `aggregation`
    (mode 'test': points -> voxels, voxels < 0 dropped, softmax(QK^T/sqrt d) V -> fc -> mean -> MLP; centroid)."""
    from fusion4landslide_b200 import nets
    z = np.load(os.path.join(golden_dir, "nets_shipped.npz"))
    model = _load(nets.ClusterFeatureNetWithAttention(), z, "agg/").to(cuda)
    p2v = _T(z["agg_p2v"], cuda)
    sptr, sidx = z["agg_spt_ptr"], _T(z["agg_spt_idx"], cuda)
    vox = p2v[sidx]
    ok = vox >= 0
    seg = torch.repeat_interleave(torch.arange(sptr.size - 1, device=cuda), _T(np.diff(sptr), cuda).long())
    cnt = torch.zeros(sptr.size - 1, dtype=torch.int64, device=cuda).index_add_(0, seg[ok], torch.ones_like(seg[ok]))
    vptr = torch.zeros(sptr.size, dtype=torch.int32, device=cuda)
    vptr[1:] = torch.cumsum(cnt, 0).to(torch.int32)
    feats = _T(z["agg_feats"], cuda)[vox[ok]].contiguous()
    coords = _T(z["agg_coords"], cuda)[vox[ok]].contiguous()
    with torch.no_grad():
        f, c = model.aggregate_segments(feats, coords, vptr)
    np.testing.assert_allclose(f.cpu().numpy(), z["agg_out_feat"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(c.cpu().numpy(), z["agg_out_coord"], atol=1e-5, rtol=1e-6)
