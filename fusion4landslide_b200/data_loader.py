"""Host mirror of src/data_loader.py (the DIPs patch front-end).

`Preprocess_Dataset` keeps the reference's constructor, `__len__` and `__getitem__` (one batch of
`points_per_batch` patches, shape (points_per_batch, 3, num_points) f32), so the loops in src/f2s3.py:104-134 and
base.py:1981-2034 run unchanged -- minus the DataLoader workers: a batch is one kernel launch and the tensor
is already on the device.  `data` / `data_overlap` may be Open3D point clouds (anything with `.points`), numpy
arrays or tensors of shape (n,3).
"""
import numpy as np
import torch

from . import ops
from ._lib import F4LError

DIPS_CAP = 1408
DIPS_CAP_LARGE = 8192


def _cloud(x, device):
    if hasattr(x, "points"):
        x = np.asarray(x.points)
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    return x.to(device=device, dtype=torch.float64).contiguous()


class Preprocess_Dataset(torch.utils.data.Dataset):
    def __init__(self, data, data_overlap, points_per_batch, feature_radius, num_points=256, device="cuda:0",
                 seed=0):
        self.device = torch.device(device)
        self.data = _cloud(data, self.device)
        self.data_overlap = self.data if data_overlap is data else _cloud(data_overlap, self.device)
        self.points_per_batch = int(points_per_batch)
        self.feature_radius = float(feature_radius)
        self.num_points = int(num_points)
        self.seed = int(seed)
        self.pcd_tree = ops.DipsIndex(self.data_overlap, self.feature_radius)       # data_loader.py:26
        self.cnt = 0

    def patches(self, offset, n, ranks=None, want_lrf=False):
        q = self.data[offset:offset + n]
        out = ops.dips_patches(self.pcd_tree, q, self.num_points, ranks=ranks,
                               seed=self.seed * 0x9E3779B97F4A7C15 + offset, want_lrf=want_lrf)
        return out

    def __getitem__(self, idx):
        offset = idx * self.points_per_batch                                           # data_loader.py:33
        if idx < 0 or offset >= self.data.shape[0]:
            raise IndexError(idx)
        patches, count = self.patches(offset, self.points_per_batch)[:2]
        cmax = int(count.max())
        if cmax > DIPS_CAP:
            # denser neighbourhoods than the fast kernel keeps on chip: those queries again through the large variant
            if cmax > DIPS_CAP_LARGE:
                raise F4LError("a point has %d neighbours within the feature radius (supported: %d); the reference's "
                               "radius rule sqrt(3)*10*resolution yields about 940" % (cmax, DIPS_CAP_LARGE))
            big = torch.nonzero(count > DIPS_CAP).reshape(-1)
            q = self.data[offset:offset + self.points_per_batch][big].contiguous()
            p2, _ = ops.dips_patches(self.pcd_tree, q, self.num_points, seed=self.seed * 0x9E3779B97F4A7C15 + offset + 1,
                                     large=True)
            patches[big] = p2
        return patches

    def __len__(self):
        return int(np.ceil(self.data.shape[0] / self.points_per_batch))               # data_loader.py:108
