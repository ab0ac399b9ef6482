#!/usr/bin/env python
"""One fusion_3d tile (3 superpoint levels) through Coarse2Fine(cfg).implement_c2f_matching() -- the profiling target for
the C3 kernels (attention pooling, descriptor NN at tile scale).  python tools/run_c3_tile.py [n_points] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fusion4landslide_b200 import configs, nets, synth  # noqa: E402
from fusion4landslide_b200.entry_c2f import Coarse2Fine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
z = np.load(os.path.join(ROOT, "tests", "golden", "nets_shipped.npz"))
model = nets.ClusterFeatureNetWithAttention()
model.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("agg/")})
model = model.to(dev).eval()
d = synth.make_scene(n, seed=1, device=dev, desc_dim=64)
tt = dict(src_pts=d["src"], tgt_pts=d["tgt"], partition_src=[d["labels_src"][k] for k in (1, 2, 3)],
          partition_tgt=[d["labels_tgt"][k] for k in (1, 2, 3)], feat_raw_src=d["src_feat"], feat_raw_tgt=d["tgt_feat"])
for _ in range(reps):
    c = Coarse2Fine(configs.fusion_config(tt, levels=[1, 2, 3], feat_aggregate_model=model))
    c.implement_c2f_matching()
torch.cuda.synchronize()
print("rows", int(c.data_output.corres_3d_refine_apply_icp.shape[0]))
