"""TEST INFRASTRUCTURE ONLY -- CPU restatement of Open3D's PointCloud.voxel_down_sample as the reference calls it
(base.py:1024-1025, :906-907).

PARITY UNPINNED: Open3D (open3d==0.19.0, requirements.txt:1) is a third-party wheel absent from this image; the
algorithm below follows its published source (cpp/open3d/geometry/PointCloud.cpp, VoxelDownSample):
voxel_min_bound = min_bound - voxel_size * 0.5, voxel index = floor((p - voxel_min_bound) / voxel_size), one output
point per occupied voxel = the mean of its points accumulated in fp64 in point order.  Open3D emits the voxels in the
iteration order of a std::unordered_map (implementation-defined); here -- and in the kernel -- they are sorted by
(ix, iy, iz).
"""
import numpy as np


def voxel_down_sample(points, voxel_size):
    """Returns centroids (V,3) f64 in ascending (ix,iy,iz) order and voxel_of_point (n,) int64."""
    pts = np.asarray(points, np.float64)
    if voxel_size <= 0:
        raise ValueError("voxel_size <= 0")
    if pts.shape[0] == 0:
        return np.zeros((0, 3)), np.zeros((0,), np.int64)
    vmb = pts.min(0) - voxel_size * 0.5
    idx = np.floor((pts - vmb) / voxel_size).astype(np.int64)
    uniq, inv = np.unique(idx, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    sums = np.zeros((uniq.shape[0], 3))
    np.add.at(sums, inv, pts)                    # unbuffered: adds in point order, like the sequential C++ loop
    cnt = np.bincount(inv, minlength=uniq.shape[0]).astype(np.float64)
    return sums / cnt[:, None], inv
