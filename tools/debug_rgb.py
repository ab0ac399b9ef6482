import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_rgb_guided_gpu import _case
from fusion4landslide_b200 import rgb_guided
from oracle import paths as opaths
cuda = torch.device("cuda:0")
corr, valid, _, patches = _case()
r = rgb_guided.local_rigid_refinement_batched(torch.from_numpy(corr).to(cuda), torch.from_numpy(valid).to(cuda),
                                              [torch.from_numpy(p) for p in patches], icp_thres=0.1)
keep_o, rows_o, per = opaths.rgb_local_rigid_refinement(corr, valid, patches, icp_thres=0.1)
rows = r["corres_3d_refine_apply_icp"].cpu().numpy(); ptr = r["seg_ptr"].cpu().numpy()
it = r["iters"].cpu().numpy(); fit = r["fitness"].cpu().numpy(); T0 = r["T_initial"].cpu().numpy(); T = r["T_icp"].cpu().numpy()
for q, p in enumerate(per):
    if p is None: continue
    a, b = ptr[q], ptr[q + 1]
    e = np.abs(rows[a:b] - rows_o[a:b]).max()
    if e > 1e-4:
        print("patch", q, "n", p["n"], "kept", p["kept"], "iters gpu/oracle", it[q], p["iters"], "fitness %.4f/%.4f" % (fit[q], p["fitness"]), "err %.3f" % e,
              "|dT0| %.2e" % np.abs(T0[q] - p["T0"]).max(), "|dT| %.2e" % np.abs(T[q] - p["T"]).max())

from fusion4landslide_b200 import ops
from oracle import icp as oicp
q = 119
a, b = ptr[q], ptr[q + 1]
rws = r["rows"].cpu().numpy()[a:b]
c = corr[rws]
src, tgt = torch.from_numpy(c[:, :3].copy()).to(cuda), torch.from_numpy(c[:, 3:6].copy()).to(cuda)
p2 = torch.tensor([0, c.shape[0]], dtype=torch.int32, device=cuda)
T0q = torch.from_numpy(per[q]["T0"]).to(cuda).reshape(1, 16).contiguous()
print("tgt extent", c[:, 3:6].min(0), c[:, 3:6].max(0), "src extent", c[:, :3].min(0), c[:, :3].max(0))
for mi in (0, 1, 2, 3, 5, 30):
    T64, fitg, rm, itg = ops.patch_icp(src, tgt, p2, p2, T0=T0q, max_corr_dist=0.1, max_iter=mi)
    o = oicp.icp_point_to_point(c[:, :3], c[:, 3:6], per[q]["T0"], 0.1, mi)
    print("max_iter %2d: gpu fitness %.4f rmse %.5f iters %d | oracle fitness %.4f rmse %.5f iters %d | |dT| %.2e" %
          (mi, fitg.item(), rm.item(), itg.item(), o["fitness"], o["inlier_rmse"], o["iters"],
           np.abs(T64.cpu().numpy().reshape(4, 4) - o["transformation"]).max()))
