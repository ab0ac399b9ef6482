"""Debug helper: the fusion-mode class test, printing the pairs whose dense rows deviate from the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import configs, synth, nets
from fusion4landslide_b200.entry_c2f import Coarse2Fine
from oracle import paths as opaths, fine_matching as ofm, icp as oicp

cuda = torch.device("cuda:0")
z = np.load(os.path.join(ROOT, "tests/golden/nets_shipped.npz"))
model = nets.ClusterFeatureNetWithAttention()
model.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("agg/")})
model = model.to(cuda).eval()
w = {k[4:]: z[k] for k in z.files if k.startswith("agg/")}
mode = "fusion"
d = synth.make_scene(36_000, seed=5, desc_dim=64, frac_2d=0.06)
levels_s = [d["labels_src"][k] for k in (1, 2, 3)]
levels_t = [d["labels_tgt"][k] for k in (1, 2, 3)]
tt = dict(src_pts=d["src"], tgt_pts=d["tgt"], partition_src=levels_s, partition_tgt=levels_t,
          feat_raw_src=d["src_feat"], feat_raw_tgt=d["tgt_feat"], corres_3d_from_2d_idx=d["corr2d"])
c = Coarse2Fine(configs.fusion_config(tt, mode=mode, levels=[1, 2, 3], feat_aggregate_model=model))
c.implement_c2f_matching()
torch.cuda.synchronize()
di = c.data_interim
v2p = {k: di["idx_voxel2pts_" + k].cpu().numpy() for k in ("src", "tgt")}
o = opaths.c2f_tile(d["src"].numpy(), d["tgt"].numpy(), [x.numpy() for x in levels_s], [x.numpy() for x in levels_t],
                    d["src_feat"].numpy(), d["tgt_feat"].numpy(), w, voxel_size=float(c.method.voxel_size),
                    corr2d=d["corr2d"].numpy(), coarse=mode, fine=mode,
                    median_max_resolution=float(np.float32(c.para.median_max_resolution)), v2p_given=v2p)
for lv in range(3):
    fr = c.fine_results_multiple[lv]
    of = o["levels"][lv]["fine"]
    K, st, it = fr.K.cpu().numpy(), fr.status.cpu().numpy(), fr.iters.cpu().numpy()
    fit, rm = fr.fitness.cpu().numpy(), fr.rmse.cpu().numpy()
    T = fr.T.cpu().numpy()
    print("level", lv, "pairs", len(K), "K equal", (K == of["K"]).all(), "status equal", (st == of["status"]).all(),
          "iters differ", int((it != of["iters"]).sum()))
    for q in np.nonzero((it != of["iters"]) | (st != of["status"]))[0]:
        dT = np.abs(T[q] - of["T"][q]).max()
        print("   pair %d K=%d status %d/%d iters %d/%d fitness %.4f/%.4f rmse %.5f/%.5f |dT| %.3e" %
              (q, K[q], st[q], of["status"][q], it[q], of["iters"][q], fit[q], of["fitness"][q], rm[q], of["rmse"][q], dT))
        if dT > 1e-3:
            corr = of["corr"][q]
            A = d["src"].numpy()[corr[:, 0]]; B = d["tgt"].numpy()[corr[:, 1]]
            print("      Tsvd oracle\n", of["Tsvd64"][q])
            print("      T gpu\n", T[q], "\n      T oracle\n", of["T"][q])
            # re-run the oracle ICP with a trace of fitness per iteration
            Tcur = of["Tsvd64"][q].copy()
            for itn in range(8):
                r = oicp.icp_point_to_point(A, B, Tcur, 0.1, 1)
                print("        it", itn, "fitness %.4f rmse %.6f" % (r["fitness"], r["inlier_rmse"]))
                Tcur = r["transformation"]

do = c.data_output
for lv in range(3):
    fr = c.fine_results_multiple[lv]
    of = o["levels"][lv]["fine"]
    dn = do.corres_3d_refine_apply_icp_multiple[lv].cpu().numpy()
    od = ofm.stack(of["dense"])
    sizes = [x.shape[0] for x in of["dense"] if x is not None]
    qs = [q for q, x in enumerate(of["dense"]) if x is not None]
    ends = np.cumsum(sizes)
    T = fr.T.cpu().numpy()
    for q, a, b in zip(qs, ends - np.array(sizes), ends):
        e = np.abs(dn[a:b] - od[a:b]).max()
        if e > 1e-4:
            print("level", lv, "pair", q, "rows", a, b, "max row diff %.4f" % e, "K", of["K"][q], "iters", of["iters"][q], fr.iters[q].item(),
                  "|dT| %.3e" % np.abs(T[q] - of["T"][q]).max(), "fitness %.4f/%.4f" % (fr.fitness[q].item(), of["fitness"][q]))
            print("   T gpu\n", T[q], "\n   T oracle\n", of["T"][q], "\n   Tsvd oracle\n", of["Tsvd64"][q])
            print("   T64 gpu\n", fr.T64[q].cpu().numpy())
