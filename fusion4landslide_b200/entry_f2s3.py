"""Class-level entry point of the F2S3 method: `Deformation_Analyze(config, src, tgt)` with the four stage methods
main_f2s3.py:72-81 calls (compute_features, implement_segmentation, correspondence_searching,
correspondence_pruning), state in the same attributes (`.correspondences`, `.svl_type`, `.src_tile_feat`, ...).

  HotPathMixin     `_compute_median_resolution` (A1, src/f2s3.py:481-508), `correspondence_searching` (B1, :248-298),
                   `correspondence_pruning` (F4 F1 A3, :321-479) as batched launches: ONE pass of the filtering network
                   over all supervoxels (nets.FilteringNetwork.compute_weights_segments), one launch for the
                   Kabsch -> median -> refit tail, the magnitude gates and the optional C2C gap filling.
  StandaloneBase   constructor / feature + segmentation loaders for use without the reference tree.
`bind(base)` = class Deformation_Analyze(HotPathMixin, base).
"""
import os
import os.path as osp

import numpy as np
import torch

from . import f2s3 as hot
from . import ops
from .entry_c2f import _Cloud, _points_of
from .functions import compute_c2c

I32 = torch.int32


def _dev(self):
    d = getattr(self, "device", None)
    if isinstance(d, str):
        d = torch.device(d)
    if d is None or d.type != "cuda":
        d = torch.device("cuda", torch.cuda.current_device())
    return d


class HotPathMixin:
    def _info(self, msg):
        if getattr(self, "verbose", False) and getattr(self, "logging", None) is not None:
            self.logging.info(msg)

    # ---- A1 ---------------------------------------------------------------------------------------------------
    def _compute_median_resolution(self):
        """src/f2s3.py:481-508: k=2 self query of both epochs, max of the two medians (float)."""
        dev = _dev(self)
        s = torch.from_numpy(_points_of(self.src_tile_non_overlap_pcd)).to(dev)
        t = torch.from_numpy(_points_of(self.tgt_tile_non_overlap_pcd)).to(dev)
        piv = torch.minimum(s.min(0).values, t.min(0).values)          # f32 kernels: tile-local coordinates
        return float(ops.median_resolution((s - piv).float().contiguous(), (t - piv).float().contiguous()).item())

    # ---- B1 ---------------------------------------------------------------------------------------------------
    def correspondence_searching(self):
        """src/f2s3.py:248-298.  The hnswlib index (approximate) becomes the exact tensor-core search; rows
        [src_xyz | tgt_xyz[label]] float64 as in the reference."""
        if not self.config.correspondence_searching:
            if not osp.exists(self.corr_path):
                if osp.exists(self.corr_path.replace('.npz', '.txt')):
                    self.corr_path = self.corr_path.replace('.npz', '.txt')
                else:
                    raise FileNotFoundError(f"The correspondence path '{self.corr_path}' is not found")
            self._load_correspondences()
            self._info('Skip correspondence searching. Load it from path')
            return
        self._info('Start correspondence searching ...')
        dev = _dev(self)
        fs = self.src_tile_feat.to(dev, torch.float32).contiguous()
        ft = self.tgt_tile_feat.to(dev, torch.float32).contiguous()
        labels, _ = ops.desc_nn(fs, ft)                                # :273-281
        self.labels = labels
        lab = labels.long().cpu().numpy()
        src = _points_of(self.src_tile_non_overlap_pcd)
        tgt = _points_of(self.tgt_tile_non_overlap_pcd)
        self.correspondences = np.concatenate((src, tgt[lab, :]), axis=1)           # :284-285
        if getattr(self, "write_interim_files", True):                 # :286-294
            mag = np.linalg.norm(self.correspondences[:, :3] - self.correspondences[:, 3:6], axis=1)
            interim = np.hstack((self.correspondences[:, :3], mag[:, None]))
            os.makedirs(osp.join(self.output_path, 'results'), exist_ok=True)
            np.savetxt(osp.join(self.output_path, 'results', f'f2s3_dvfms_without_pruning_of_tile_{self.tile_id}.txt'), interim)
            interim[0, 3] = 0
            interim[1, 3] = 5
            np.savetxt(osp.join(self.output_path, 'results',
                                f'f2s3_dvfms_without_pruning_of_tile_{self.tile_id}_visualize_0_5.txt'), interim)
        if self.save_interim:
            os.makedirs(osp.dirname(self.corr_path), exist_ok=True)
            np.savez(self.corr_path, corr=self.correspondences)        # :296-298

    def _load_correspondences(self):                                   # :300-319
        corr = np.load(self.corr_path) if osp.splitext(self.corr_path)[1] == '.npz' else np.loadtxt(self.corr_path)
        self.correspondences = torch.as_tensor(np.asarray(corr['corr'] if hasattr(corr, 'files') else corr)).to(_dev(self))

    # ---- F4 F1 A3 ---------------------------------------------------------------------------------------------
    def _batched_filter_net(self):
        """The filtering network in its all-supervoxels form.  A reference `FilteringNetwork` (same parameters,
        per-supervoxel forward) is re-hosted in nets.FilteringNetwork."""
        from . import nets
        net = self._outlier_removal_nn()
        if net is None:
            raise RuntimeError("correspondence_pruning needs config.outlier_removal_nn (outlier_removal: True)")
        if hasattr(net, "compute_weights_segments"):
            return net
        cached = getattr(self, "_b200_filter_net", None)
        if cached is None or cached[0] is not net:
            mine = nets.FilteringNetwork().to(_dev(self))
            mine.load_state_dict(net.state_dict())
            mine.eval()
            self._b200_filter_net = cached = (net, mine)
        return cached[1]

    def correspondence_pruning(self):
        """src/f2s3.py:321-479 with the per-supervoxel loop (:340-366) as one batched pass."""
        self._info('Start correspondence pruning ...')
        dev = _dev(self)
        cfg = self.config
        corr_all = self.correspondences
        if torch.is_tensor(corr_all):
            corr_all = corr_all.detach().cpu().numpy()
        corr_all = np.asarray(corr_all)
        sizes = np.fromiter((len(s) for s in self.svl_type), dtype=np.int64, count=len(self.svl_type))
        order = np.concatenate([np.asarray(s, dtype=np.int64) for s in self.svl_type]) if len(self.svl_type) else np.zeros(0, np.int64)
        seg_ptr = torch.zeros(sizes.size + 1, dtype=I32)
        seg_ptr[1:] = torch.from_numpy(np.cumsum(sizes)).to(I32)
        seg_ptr = seg_ptr.to(dev)
        save_coords = corr_all[order]                                  # rows in supervoxel order (:366,:376)
        X64 = torch.from_numpy(np.ascontiguousarray(save_coords, dtype=np.float64)).to(dev)
        with torch.no_grad():
            scaled = ops.segment_scale_maxabs(X64, seg_ptr)            # :343, f64 division then .float() (:346)
            scores = self._batched_filter_net().compute_weights_segments(scaled, seg_ptr, scale=False)
        self.scores = scores
        coeff = 2.5 if 'Rockfall_Simulator' in cfg.data_dir else 1.0
        R, t, robust, _ = hot.filter_input_tail(X64.float().contiguous(), scores, seg_ptr, coeff)   # outlier_classifier.py:71-105
        self.rot_est, self.trans_est, self.robust_estimate = R, t, robust
        seg_of_row = torch.repeat_interleave(torch.arange(sizes.size, device=dev), (seg_ptr[1:] - seg_ptr[:-1]).long())
        keep = scores > 0.99999                                        # :363
        if cfg.refine_results:
            keep = keep | robust[seg_of_row]                           # :351-360 (saved rows stay unrefined, quirk q6)
        inlier_idx = torch.nonzero(keep).reshape(-1).cpu().numpy()     # :372-374
        filtered_results = save_coords[inlier_idx, :]                  # :380-381 (float64 numpy, like the reference)
        filtered_magnitudes = np.linalg.norm(filtered_results[:, 3:6] - filtered_results[:, 0:3], axis=1)
        self._info('{} points out of {} were classified as inlier'.format(filtered_results.shape[0], save_coords.shape[0]))
        out_dir = osp.join(self.output_path, 'results')
        os.makedirs(out_dir, exist_ok=True)
        final_results = np.concatenate((filtered_results, filtered_magnitudes.reshape(-1, 1)), axis=1)
        final_results = final_results[final_results[:, 6] <= cfg.max_disp_magnitude]             # :392-393 (non-strict)
        self.final_results = final_results
        write = getattr(self, "write_results", True)
        if write:
            np.savetxt(osp.join(out_dir, 'f2s3_dvfs_of_tile_{}.txt'.format(self.tile_id)), final_results[:, :6])
            np.savetxt(osp.join(out_dir, 'f2s3_dvfms_of_tile_{}.txt'.format(self.tile_id)), final_results[:, [0, 1, 2, 6]])
            if final_results.shape[0] > 2:                             # :399-403
                vis = final_results.copy()
                vis[0, 6], vis[1, 6] = 0, 5
                np.savetxt(osp.join(out_dir, f'f2s3_dvfms_of_tile_{self.tile_id}_visualize_0_5.txt'), vis[:, [0, 1, 2, 6]])
        if cfg.max_disp_magnitude > 0:                                 # :419-424 (strict)
            sel = np.where(filtered_magnitudes < cfg.max_disp_magnitude)[0].reshape(-1)
            filtered_results, filtered_magnitudes, inlier_idx = filtered_results[sel, :], filtered_magnitudes[sel], inlier_idx[sel].reshape(-1)
        if cfg.filter_median_magnitude:                                # :427-463
            median_mag = np.median(filtered_magnitudes)
            mag_inlier = np.where(filtered_magnitudes < 30 * median_mag)[0]
            filtered_results, filtered_magnitudes = filtered_results[mag_inlier, :], filtered_magnitudes[mag_inlier]
            if write:
                d = osp.join(out_dir, 'filtered_by_magnitude')
                os.makedirs(d, exist_ok=True)
                np.savetxt(osp.join(d, 'f2s3_dvfms_filtered_by_median_mag_of_tile_{}.txt'.format(self.tile_id)),
                           np.concatenate((filtered_results[:, :3], filtered_magnitudes.reshape(-1, 1)), axis=1))
            if cfg.fill_gaps_c2c:
                self._fill_gaps(save_coords, inlier_idx[mag_inlier], filtered_magnitudes, out_dir, write)
        elif cfg.fill_gaps_c2c:                                        # :466-477
            self._fill_gaps(save_coords, inlier_idx, filtered_magnitudes, out_dir, write)
        self.filtered_results, self.filtered_magnitudes = filtered_results, filtered_magnitudes
        return None

    def _fill_gaps(self, save_coords, idx, mags, out_dir, write):
        c2c = compute_c2c(save_coords[:, 0:3], _points_of(self.tgt_tile_non_overlap_pcd)).reshape(-1)   # A3
        c2c[idx] = mags
        self.c2c_displacements = c2c
        if write:
            d = osp.join(out_dir, 'combined_with_c2c')
            os.makedirs(d, exist_ok=True)
            np.savetxt(osp.join(d, 'f2s3_dvfms_combined_with_c2c_of_tile_{}.txt'.format(self.tile_id)),
                       np.concatenate((save_coords[:, 0:3], c2c.reshape(-1, 1)), axis=1))


class StandaloneBase(object):
    """Constructor + loaders of Deformation_Analyze (src/f2s3.py:19-246) for use without the reference tree.
    Feature networks and the native supervoxel segmentation stay in the reference; their result files are read."""

    def __init__(self, config, src_tile_overlap_path, tgt_tile_overlap_path):
        from .piecewise_icp import _read_xyz
        self.config = config
        self.logging = config.get("logging", None)
        self.verbose = config.get("verbose", False)
        self.voxel_size = config.get("voxel_size", 0.1)
        self.points_per_batch = config.get("points_per_batch", 1000)
        self.device = config.get("device", "cuda")
        self.batch_size = config.get("batch_size", 1)
        self.num_workers = config.get("num_workers", 0)
        # in-memory tile: config.tile_tensors = dict(src_pts, tgt_pts (n,3) float64 numpy / tensors, src_feat, tgt_feat
        # (n,D), svl_idx (n,) supervoxel label per source point); the two path arguments are then unused
        self.tile_tensors = config.get("tile_tensors", None)
        if self.tile_tensors is not None:
            self.src_tile_overlap_pcd = _Cloud(self.tile_tensors["src_pts"])
            self.tgt_tile_overlap_pcd = _Cloud(self.tile_tensors["tgt_pts"])
        else:
            self.src_tile_overlap_pcd = _Cloud(_read_xyz(src_tile_overlap_path))
            self.tgt_tile_overlap_pcd = _Cloud(_read_xyz(tgt_tile_overlap_path))
        self.src_tile_non_overlap_path, self.tgt_tile_non_overlap_path = src_tile_overlap_path, tgt_tile_overlap_path
        self.src_tile_non_overlap_pcd, self.tgt_tile_non_overlap_pcd = self.src_tile_overlap_pcd, self.tgt_tile_overlap_pcd
        self.tile_id = config.tile_id
        self.output_path = osp.join(config.output_dir, config.output_folder)
        self.src_tile_feat = self.tgt_tile_feat = self.correspondences = None
        self.feat_compute = config.feat_compute
        self.pcd_segment = config.pcd_segment
        self.outlier_removal = config.outlier_removal
        self.sv_type = None
        self.segment_type = config.segment_type
        self.save_interim = config.save_interim
        self.small_patch_removal = config.small_patch_removal
        self.corr_path = osp.join(self.output_path, 'correspondences', f'corr_tile_{self.tile_id}.npz')
        self.write_interim_files = config.get("write_interim_files", True)
        self.write_results = config.get("write_results", True)

    @property
    def _feat_path(self):
        return osp.join(self.output_path, 'features', f'features_tile_{self.tile_id}.npz')

    @property
    def _segment_path(self):
        if self.segment_type == 'supervoxel':
            folder = 'svl_segment'
        elif self.segment_type == 'superpoint':
            folder = osp.join('spt_segment', self.config.spt_color_level)
        else:
            raise NotImplementedError
        return osp.join(self.output_path, folder, f'segment_tile_{self.tile_id}.txt')

    def _feat_desc_nn(self, x):
        if self.feat_compute:
            return self.config.feat_desc_nn(x)

    def _outlier_removal_nn(self):
        if self.outlier_removal:
            return self.config.outlier_removal_nn

    def compute_features(self):                                        # src/f2s3.py:91-164
        dev = _dev(self)
        if self.tile_tensors is not None and "src_feat" in self.tile_tensors:
            self.src_tile_feat = torch.as_tensor(self.tile_tensors["src_feat"]).to(dev)
            self.tgt_tile_feat = torch.as_tensor(self.tile_tensors["tgt_feat"]).to(dev)
            return None
        if not self.config.feat_compute:
            if not osp.exists(self._feat_path):
                raise FileNotFoundError(f"The feature path '{self._feat_path}' is not found")
            f = np.load(self._feat_path)
            self.src_tile_feat = torch.from_numpy(f['src_feat']).to(dev)
            self.tgt_tile_feat = torch.from_numpy(f['tgt_feat']).to(dev)
            return None
        if self.config.feat_type != 'DIPs':
            raise NotImplementedError
        from .data_loader import Preprocess_Dataset
        radius = np.sqrt(3) * (10 * self._compute_median_resolution())  # :106
        out = []
        for a, b in ((self.src_tile_non_overlap_pcd, self.src_tile_overlap_pcd),
                     (self.tgt_tile_non_overlap_pcd, self.tgt_tile_overlap_pcd)):
            ds = Preprocess_Dataset(a, b, self.points_per_batch, radius, device=str(dev))
            with torch.no_grad():
                out.append(torch.cat([self._feat_desc_nn(ds[i])[0] for i in range(len(ds))], dim=0))
        self.src_tile_feat, self.tgt_tile_feat = out
        if self.save_interim:
            os.makedirs(osp.dirname(self._feat_path), exist_ok=True)
            np.savez_compressed(self._feat_path, src_feat=self.src_tile_feat.cpu(), tgt_feat=self.tgt_tile_feat.cpu())
        return None

    def implement_segmentation(self):                                  # src/f2s3.py:166-238
        if self.tile_tensors is not None and "svl_idx" in self.tile_tensors:
            lab = self.tile_tensors["svl_idx"]
            lab = lab.cpu().numpy() if torch.is_tensor(lab) else np.asarray(lab)
            self.svl_type = supervoxel_lists(lab, 10 if self.small_patch_removal else 1, _dev(self))
            return
        if self.pcd_segment:
            raise NotImplementedError("the native supervoxel / superpoint segmentation stays in the reference; "
                                      "set pcd_segment: False to load its result file")
        if not osp.exists(self._segment_path):
            raise FileNotFoundError(f"The segmentation result path '{self._segment_path}' is not found")
        seg = np.loadtxt(self._segment_path)
        svl_idx = seg[:, -2] if self.segment_type == 'superpoint' else seg[:, -1]
        self.svl_type = supervoxel_lists(svl_idx, 10 if self.small_patch_removal else 1, _dev(self))


def supervoxel_lists(svl_idx, min_count, device):
    """src/f2s3.py:213-237: index arrays of the supervoxels with more than `min_count` points, in ascending label
    order (np.unique), points ascending -- one labels->CSR launch instead of a boolean mask per label."""
    lab = torch.from_numpy(np.asarray(svl_idx).reshape(-1).astype(np.int64)).to(device)
    _, ptr, idx, _ = ops.labels_to_csr(lab.contiguous(), int(min_count))
    p = ptr.cpu().numpy()
    return np.split(idx.cpu().numpy().astype(np.int64), p[1:-1]) if p.size > 1 else []


def bind(base):
    class Deformation_Analyze(HotPathMixin, base):
        pass
    Deformation_Analyze.__doc__ = "Drop-in for `from src.f2s3 import Deformation_Analyze` (main_f2s3.py:12,72-81)."
    return Deformation_Analyze


Deformation_Analyze = bind(StandaloneBase)
