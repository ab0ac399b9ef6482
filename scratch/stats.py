import torch, numpy as np, sys
sys.path.insert(0, '.')
from fusion4landslide_b200 import pipeline, synth
dev = torch.device('cuda:0')
d = synth.make_scene(781250, seed=0, device=dev) if "--v1" not in sys.argv else synth.make_tile(781250, seed=0, device=dev, patch_pts=256)
t = pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"])
r, med = pipeline.displacement_field(t)
torch.cuda.synchronize()
it = r.iters.cpu().numpy(); K = r.K.cpu().numpy(); st = r.status.cpu().numpy()
print('pairs', len(it), 'status counts', np.bincount(st, minlength=3), 'K mean/max', K.mean(), K.max())
print('iters mean', it[st == 0].mean(), 'hist', np.bincount(it[st == 0], minlength=31))
print('fitness mean', r.fitness.cpu().numpy()[st == 0].mean(), 'rmse mean', r.rmse.cpu().numpy()[st == 0].mean())
print('n_src_items', t.n_src_items, 'counts', r.counts.tolist(), 'med', med.item())
import ctypes
from fusion4landslide_b200 import _lib
L = _lib.lib()
if hasattr(L, 'f4l_debug_counters'):
    buf = (ctypes.c_ulonglong * 24)()
    L.f4l_debug_counters(buf, 1)
    r, med = pipeline.displacement_field(t)
    torch.cuda.synchronize()
    L.f4l_debug_counters(buf, 1)
    v = list(buf)
    print('scans', v[0], 'exact', v[1], 'iterations(matches)', v[2], 'points*iters', v[3], 'scan frac', v[0] / max(v[3], 1), 'mean moved um', v[4] / max(v[2], 1))
    names = {17: 'rigidity staging', 8: 'rigidity', 9: 'procrustes', 10: 'icp stage', 11: 'phase1 keep-check', 12: 'lane scans', 13: 'exact scans', 14: 'phase3 moments', 15: 'update svd', 16: 'total'}
    tot = max(v[16], 1)
    for k_, n_ in names.items():
        print('  %-20s %8.1f kcycles/patch  %5.1f%%' % (n_, v[k_] / 3136 / 1e3, 100.0 * v[k_] / tot))
