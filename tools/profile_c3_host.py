#!/usr/bin/env python
"""Host-side profile of one C3 tile through the class-level entry point (cProfile + per-stage wall time with syncs)."""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import configs, nets, synth
from fusion4landslide_b200.entry_c2f import Coarse2Fine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 625_000
dev = torch.device("cuda:0")
z = np.load(os.path.join(ROOT, "tests", "golden", "nets_shipped.npz"))
model = nets.ClusterFeatureNetWithAttention()
model.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("agg/")})
model = model.to(dev).eval()
d = synth.make_scene(n, seed=1, device=dev, desc_dim=64)
tt = dict(src_pts=d["src"], tgt_pts=d["tgt"], partition_src=[d["labels_src"][k] for k in (1, 2, 3)],
          partition_tgt=[d["labels_tgt"][k] for k in (1, 2, 3)], feat_raw_src=d["src_feat"], feat_raw_tgt=d["tgt_feat"])
def run():
    c = Coarse2Fine(configs.fusion_config(tt, levels=[1, 2, 3], feat_aggregate_model=model))
    c.implement_c2f_matching()
    return c
run(); torch.cuda.synchronize()
t0 = time.perf_counter(); run(); torch.cuda.synchronize(); print("tile wall %.1f ms" % (1e3 * (time.perf_counter() - t0)))
# stage timing with syncs
c = Coarse2Fine(configs.fusion_config(tt, levels=[1, 2, 3], feat_aggregate_model=model))
names = ["_voxel_subsampling", "prepare_pts2spt_dict", "global_matches_from_3d", "_compute_spt_feat_and_coord_with_fused_feats",
         "coarse_matching_with_different_types", "fine_matching_with_different_types", "save_process_dvf", "compute_point_feat"]
acc = {k: 0.0 for k in names}
for k in names:
    f = getattr(c, k)
    def wrap(f=f, k=k):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); acc[k] += time.perf_counter() - t; return r
    setattr(c, k, wrap)
t0 = time.perf_counter(); c.implement_c2f_matching(); torch.cuda.synchronize(); tot = time.perf_counter() - t0
print("synced total %.1f ms" % (1e3 * tot)); [print("   %-48s %.1f ms" % (k, 1e3 * v)) for k, v in acc.items()]
pr = cProfile.Profile(); pr.enable(); run(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
from fusion4landslide_b200 import _lib
L = _lib.lib()
L.f4l_profile_reset(); L.f4l_profile_enable(1)
run(); torch.cuda.synchronize()
L.f4l_profile_enable(0)
tab = _lib.profile_table()
print("kernel table (one tile):")
for k, v in sorted(tab.items(), key=lambda kv: -kv[1][0])[:14]:
    print("   %-28s %8.3f ms  x%d" % (k, v[0], v[1]))
