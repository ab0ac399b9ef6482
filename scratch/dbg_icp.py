import numpy as np, torch, sys
sys.path.insert(0, '.')
from tests.test_icp_gpu import _tile_patches, _matched
from fusion4landslide_b200 import ops
from oracle import icp as oicp
cuda = torch.device('cuda:0')
d, (ptr_s, idx_s), (ptr_t, idx_t), m, j = _tile_patches()
a, b = m.tolist()[0], j.tolist()[0]
ps = idx_s[ptr_s[a]:ptr_s[a + 1]].numpy(); pt = idx_t[ptr_t[b]:ptr_t[b + 1]].numpy()
A, B = _matched(d, ps, pt)
print('K', A.shape)
ptr = torch.tensor([0, len(A)], dtype=torch.int32, device=cuda)
for mi in (0, 1, 2, 3, 30):
    T, fit, rmse, iters, corr = ops.patch_icp(torch.from_numpy(A).to(cuda), torch.from_numpy(B).to(cuda), ptr, ptr, max_corr_dist=0.1, max_iter=mi, want_corr=True)
    o = oicp.icp_point_to_point(A, B, None, 0.1, max_iter=mi)
    oc = -np.ones(len(A), np.int64); oc[o['correspondence_set'][:, 0]] = o['correspondence_set'][:, 1]
    c = corr.cpu().numpy()
    diff = np.nonzero(c != oc)[0]
    print('max_iter', mi, 'iters', iters.item(), o['iters'], 'fit', fit.item(), o['fitness'], 'rmse diff', rmse.item() - o['inlier_rmse'], 'Tdiff', np.abs(T.cpu().numpy()[0] - o['transformation']).max(), 'corr diffs', len(diff))
    for i in diff[:5]:
        P = A[i].astype(np.float64) @ o['transformation'][:3, :3].T + o['transformation'][:3, 3]
        dd = np.linalg.norm(B.astype(np.float64) - P, axis=1)
        print('   i', i, 'gpu', c[i], 'orc', oc[i], 'd gpu', dd[c[i]] if c[i] >= 0 else None, 'd orc', dd[oc[i]] if oc[i] >= 0 else None, 'same coords', (B[c[i]] == B[oc[i]]).all() if c[i] >= 0 and oc[i] >= 0 else None)
