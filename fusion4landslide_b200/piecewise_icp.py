"""Host mirror of src/piecewise_icp.py::Piecewise_ICP(cfg) (rows G1, A5, F5): same config keys, same
result files; the computation is one CUDA launch sequence (f4l_piecewise_icp), fp64 like Open3D / numpy.
"""
import os

import numpy as np
import torch

from . import ops


def _read_xyz(path):
    """Tile reader.  Geospatial I/O stays as in the reference (Open3D) when it is installed; .npy / .txt and
    simple PLY files (ascii or binary_little_endian with leading float/double x y z) are read directly."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        return np.load(path).astype(np.float64)[:, :3]
    if ext in (".txt", ".xyz", ".csv"):
        return np.loadtxt(path)[:, :3].astype(np.float64)
    try:
        import open3d as o3d
        return np.asarray(o3d.io.read_point_cloud(path).points, dtype=np.float64)
    except ImportError:
        pass
    with open(path, "rb") as f:
        fmt, n, props = None, 0, []
        in_vertex = False
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element"):
                in_vertex = line.split()[1] == "vertex"
                if in_vertex:
                    n = int(line.split()[2])
            elif line.startswith("property") and in_vertex:
                props.append((line.split()[1], line.split()[2]))
            elif line == "end_header":
                break
        if fmt == "ascii":
            return np.loadtxt(f, max_rows=n)[:, :3].astype(np.float64)
        if fmt != "binary_little_endian":
            raise ValueError("unsupported PLY format %r" % fmt)
        types = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
                 "char": "i1", "short": "<i2", "ushort": "<u2", "int": "<i4", "uint": "<u4", "int32": "<i4"}
        dt = np.dtype([(name, types[t]) for t, name in props])
        v = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
        return np.stack([v["x"], v["y"], v["z"]], axis=1).astype(np.float64)


def piecewise_icp(src_pts, tgt_pts, smax, number_points_min):
    """Device-level entry: (N,3) source / target points (numpy or tensor) -> dvfs (N',6) float64 cuda,
    dvfms (N',4).  Raises ValueError when no cell is unstable, like np.vstack([]) at piecewise_icp.py:197."""
    dev = torch.device("cuda", torch.cuda.current_device())
    s = torch.as_tensor(np.asarray(src_pts) if not torch.is_tensor(src_pts) else src_pts).to(dev, torch.float64).contiguous()
    t = torch.as_tensor(np.asarray(tgt_pts) if not torch.is_tensor(tgt_pts) else tgt_pts).to(dev, torch.float64).contiguous()
    dvfs, mag, counts, _ = ops.piecewise_icp(s, t, smax, number_points_min)
    c = counts.tolist()
    if c[5] == 0:
        raise ValueError("need at least one array to concatenate")
    dvfs = dvfs[:c[0]]
    return dvfs, torch.cat([dvfs[:, :3], mag[:c[0], None]], dim=1)


def Piecewise_ICP(cfg):
    """Drop-in for `from src.piecewise_icp import Piecewise_ICP` (main_piecewise_icp.py:14,93).  Reads
    cfg.{src_tile_overlap_path, tgt_tile_overlap_path, smax, number_points_min, output_root, tile_id}; cfg.threshold
    and the identity trans_init are read but unused by the reference too (piecewise_icp.py:38,41)."""
    src = _read_xyz(cfg.src_tile_overlap_path)
    tgt = _read_xyz(cfg.tgt_tile_overlap_path)
    dvfs, dvfms = piecewise_icp(src, tgt, cfg.smax, cfg.number_points_min)
    dvfs, dvfms = dvfs.cpu().numpy(), dvfms.cpu().numpy()
    out = os.path.join(cfg.output_root, "results")
    os.makedirs(out, exist_ok=True)
    tid = cfg.tile_id
    np.savetxt(os.path.join(out, "piecewise_icp_dvfms_of_tile_{}.txt".format(tid)), dvfms)       # :203-204
    np.savetxt(os.path.join(out, "piecewise_icp_dvfs_of_tile_{}.txt".format(tid)), dvfs)         # :205-206
    vis = dvfms.copy()                                                                            # :205-214 (q10)
    vis[0, 3] = 0
    vis[1, 3] = {"rockfall": 0.06, "brienz_tls": 5, "mattertal": 10}.get(getattr(cfg, "dataset", None), 10)
    np.savetxt(os.path.join(out, "piecewise_dvfms_visualize_of_tile_{}.txt".format(tid)), vis)
    return None
