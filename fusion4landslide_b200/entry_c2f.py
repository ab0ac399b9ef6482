"""Class-level entry point of the fusion method: `Coarse2Fine(cfg).implement_c2f_matching()` as called from
main_fusion.py:147-148, with the state kept in the reference's `data_input_3d / data_interim / data_output`
EasyDict fields.

Two pieces:

  HotPathMixin     the hot methods of `Coarse2Fine_Base` / `Coarse2Fine` re-expressed as batched launches of
                   libf4l_b200.so: `_compute_median_resolution` (A1), `_voxel_subsampling` (A2 + 8f-3),
                   `prepare_pts2spt_dict` (F6), `global_matches_from_3d` (B2), `coarse_matching_with_different_types`
                   (B3, B4), `fine_matching_with_different_types` (F2 F3 D2 E1 D5 A4), the level loop and the merge
                   of `implement_c2f_matching` (coarse_to_fine_matching.py:201-290, M1).
  StandaloneBase   what the hot methods need around them when the reference tree is not importable (this
                   repository's tests and benchmarks): config plumbing (`_initialize`), tile readers, partition /
                   feature loaders, the result writers -- file formats and field names as in the reference.

`bind(base)` builds `class Coarse2Fine(HotPathMixin, base)`.  In the reference's environment `base` is the
reference's own `Coarse2Fine_Base`, so 2D matching / lifting, partitioning, descriptor networks, I/O and
visualisation stay the reference's code and only the hot methods are replaced (compat/src/coarse_to_fine_matching.py).
"""
import os
import os.path as osp

import numpy as np
import torch

from . import coarse_to_fine as c2f
from . import ops, pipeline
from .functions import _dev_f32

I32 = torch.int32
I64 = torch.int64


try:                                    # the reference uses easydict; a minimal stand-in keeps this importable without it
    from easydict import EasyDict as edict
except ImportError:                     # pragma: no cover - depends on the environment
    class edict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, edict):
                v = edict(v)
            super().__setitem__(k, v)

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        __setattr__ = __setitem__


def _points_of(pcd_or_tensor):
    """float64 numpy (n,3) of an Open3D cloud / anything with `.points`, a tensor or an array."""
    if hasattr(pcd_or_tensor, "points"):
        return np.asarray(pcd_or_tensor.points, dtype=np.float64)
    if torch.is_tensor(pcd_or_tensor):
        return pcd_or_tensor.detach().cpu().numpy().astype(np.float64)
    return np.asarray(pcd_or_tensor, dtype=np.float64)


class _Cloud:
    """Stand-in for an Open3D PointCloud where only `.points` is consumed downstream (data_loader.py:25-26).
    Built from a numpy array or from a device tensor; the host copy of a device tensor is made on first use."""

    def __init__(self, points64):
        self._dev = points64 if torch.is_tensor(points64) else None
        self._host = None if torch.is_tensor(points64) else points64

    @property
    def points(self):
        if self._host is None:
            self._host = self._dev.detach().cpu().numpy().astype(np.float64)
        return self._host

    def device_points(self, dev):
        """float64 (n,3) tensor on `dev` without a host round trip when the cloud already lives there."""
        if self._dev is not None and self._dev.device.type == dev.type and dev.index in (None, self._dev.device.index):
            return self._dev.to(torch.float64)
        return torch.from_numpy(np.ascontiguousarray(self.points, dtype=np.float64)).to(dev)


class _SegmentList:
    """The reference's `idx_spt2pts_*` (a Python list with one index tensor per patch, base.py:1340-1344) as a read-only
    sequence over the CSR tables the kernels use: element i is the view idx[ptr[i]:ptr[i+1]] (int64), created on access.
    Building ~10^4 tensor objects per level and epoch up front cost more host time than the level's kernels."""
    _next_key = [0]

    def __init__(self, idx64, ptr_host):
        self.idx, self.ptr = idx64, ptr_host
        _SegmentList._next_key[0] += 1
        self.key = ("seg", _SegmentList._next_key[0])

    def __len__(self):
        return len(self.ptr) - 1

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self.idx[int(self.ptr[i]):int(self.ptr[i + 1])]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class _IndexedList:
    """`[base[a] for a in positions]` (spt_corres_src / spt_corres_tgt, base.py:3156-3157) without building it: the
    positions stay on the device until an element is asked for."""

    def __init__(self, base, positions):
        self.base, self.pos = base, positions
        self._host = None
        _SegmentList._next_key[0] += 1
        self.key = ("pairs", _SegmentList._next_key[0])

    def __len__(self):
        return int(self.pos.numel())

    def _positions(self):
        if self._host is None:
            self._host = self.pos.cpu().numpy()
        return self._host

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        return self.base[int(self._positions()[int(i)])]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _list_key(lst):
    """Identity of a list of index tensors that survives EasyDict's copy-on-assign of list values."""
    key = getattr(lst, "key", None)
    if key is not None:
        return key
    if not len(lst):
        return (0, 0, 0)
    return (len(lst), int(lst[0].data_ptr()), int(lst[-1].data_ptr()))


class _LevelTables:
    """CSR form of one partition level (what `prepare_pts2spt_dict` returns as lists) + the tables the kernels use."""
    __slots__ = ("lab_s", "ptr_s", "idx_s", "pop_s", "lab_t", "ptr_t", "idx_t", "pop_t", "list_s", "list_t")


class HotPathMixin:
    """Hot methods of Coarse2Fine_Base on the B200 kernels.  Field names: SURVEY 9.1."""

    # ---- helpers ------------------------------------------------------------------------------------------
    def _b200_dev(self):
        d = getattr(self, "device", None)
        if isinstance(d, str):
            d = torch.device(d)
        if d is None or d.type != "cuda":
            d = torch.device("cuda", torch.cuda.current_device())
        return d

    def _b200_tables(self):
        if not hasattr(self, "_b200_levels"):
            self._b200_levels = {}       # _list_key(idx_spt2pts_src) -> _LevelTables
            self._b200_pairs = {}        # _list_key(spt_corres_src) -> (m, j, tables)
        return self._b200_levels, self._b200_pairs

    def _info(self, msg):
        if getattr(self, "verbose", False) and getattr(self, "logging", None) is not None:
            self.logging.info(msg)

    # ---- A1 -------------------------------------------------------------------------------------------------
    def _compute_median_resolution(self):
        """base.py:2716-2754: k=2 self query of both (sub-sampled) epochs, max of the two medians."""
        med = c2f.compute_median_resolution(self.data_interim.src_pts_sub, self.data_interim.tgt_pts_sub)
        val = float(med.item())
        self.para.median_max_resolution = val                      # :2751
        return val

    # ---- A2 + 8(f)-3 ----------------------------------------------------------------------------------------
    def _voxel_subsampling(self):
        """base.py:1012-1057: voxel size = median resolution of the raw tile, Open3D voxel_down_sample of both
        epochs, nearest raw point of every voxel (idx_voxel2pts_*), inverse maps (idx_pts2voxel_*, -1 default)."""
        di, d3 = self.data_interim, self.data_input_3d
        dev = self._b200_dev()
        di.src_pts_sub = _dev_f32(d3.src_pts, dev)                 # :1019-1022 (the raw tile first)
        di.tgt_pts_sub = _dev_f32(d3.tgt_pts, dev)
        self.method.voxel_size = self._compute_median_resolution() # :1023
        for name in ("src", "tgt"):
            pcd = d3.get(name + "_pcd", None)
            if isinstance(pcd, _Cloud):
                raw64 = pcd.device_points(dev)
            elif pcd is not None:
                raw64 = torch.from_numpy(_points_of(pcd)).to(dev)
            elif torch.is_tensor(d3[name + "_pts"]) and d3[name + "_pts"].device == dev:
                raw64 = d3[name + "_pts"].to(torch.float64)
            else:
                raw64 = torch.from_numpy(_points_of(d3[name + "_pts"])).to(dev)
            sub64 = ops.voxel_downsample(raw64.contiguous(), self.method.voxel_size)          # :1024-1025
            di[name + "_pcd_sub"] = _Cloud(sub64)
            sub = sub64.float().contiguous()                       # pcd2tensor -> float32 (o3d_tools.py:251)
            di[name + "_pts_sub"] = sub
            v2p, p2v = c2f.voxel_subsampling_maps(sub, _dev_f32(d3[name + "_pts"], dev))      # :1038-1057
            di["idx_voxel2pts_" + name] = v2p                      # (the reference keeps these two as numpy; every
            di["idx_pts2voxel_" + name] = p2v                      #  consumer below accepts tensors)

    # ---- F6 ---------------------------------------------------------------------------------------------------
    def prepare_pts2spt_dict(self):
        """base.py:1301-1351: patches with count > num_min_matches_for_small_patch (or all of them), points of a
        patch in ascending index order, patches in ascending label order.  Besides the reference's list form the
        CSR tables are cached for the kernels."""
        self._info('Prepare point-superpoint indices...')
        di = self.data_interim
        dev = self._b200_dev()
        min_pts = int(self.method.num_min_matches_for_small_patch) if self.method.small_patch_removal else 0
        t = _LevelTables()
        t.lab_s, t.ptr_s, t.idx_s, t.pop_s = ops.labels_to_csr(di.idx_pts2spt_src.to(dev, I64).contiguous(), min_pts)
        t.lab_t, t.ptr_t, t.idx_t, t.pop_t = ops.labels_to_csr(di.idx_pts2spt_tgt.to(dev, I64).contiguous(), min_pts)
        t.list_s = _SegmentList(t.idx_s.long(), t.ptr_s.cpu().numpy())
        t.list_t = _SegmentList(t.idx_t.long(), t.ptr_t.cpu().numpy())
        di.idx_spt_src, di.idx_spt_tgt = t.lab_s, t.lab_t          # :1340-1344
        di.idx_spt2pts_src, di.idx_spt2pts_tgt = t.list_s, t.list_t
        levels, _ = self._b200_tables()
        levels[_list_key(t.list_s)] = t
        self._info('Preparing superpoint indices is done!')

    def _level_tables(self):
        """Tables of the level `data_interim.idx_spt2pts_src` currently points at (rebuilt from the lists when they
        were produced elsewhere, e.g. by the reference's own prepare_pts2spt_dict)."""
        levels, _ = self._b200_tables()
        di = self.data_interim
        t = levels.get(_list_key(di.idx_spt2pts_src))
        if t is not None:
            return t
        dev = self._b200_dev()
        t = _LevelTables()

        def pack(lst, n_pts):
            cnt = torch.tensor([0] + [int(x.numel()) for x in lst], dtype=I64)
            p = torch.cumsum(cnt, 0).to(dev, I32)
            idx = (torch.cat([x.reshape(-1) for x in lst]) if lst else torch.empty(0, dtype=I64)).to(dev, I32).contiguous()
            pop = torch.full((n_pts,), -1, dtype=I32, device=dev)
            if idx.numel():
                seg = torch.repeat_interleave(torch.arange(len(lst), device=dev, dtype=I32), (p[1:] - p[:-1]).long())
                pop[idx.long()] = seg
            return p, idx, pop

        t.ptr_s, t.idx_s, t.pop_s = pack(di.idx_spt2pts_src, self.data_input_3d.src_pts.shape[0])
        t.ptr_t, t.idx_t, t.pop_t = pack(di.idx_spt2pts_tgt, self.data_input_3d.tgt_pts.shape[0])
        t.lab_s, t.lab_t = torch.as_tensor(di.idx_spt_src).to(dev), torch.as_tensor(di.idx_spt_tgt).to(dev)
        t.list_s, t.list_t = list(di.idx_spt2pts_src), list(di.idx_spt2pts_tgt)
        levels[_list_key(t.list_s)] = t
        return t

    # ---- B2 ---------------------------------------------------------------------------------------------------
    def global_matches_from_3d(self):
        """base.py:2756-2923.  Every `global_matching_from_3d_type` of the reference ('hnsw', 'cdist', 'cdist_cpu',
        'faiss') maps to the exact tensor-core search (= its 'cdist' branches; the HNSW indexes are approximate)."""
        self._info('Start global matches from 3d...')
        st = self.method.global_matching_from_3d_type
        if st not in ('hnsw', 'cdist_cpu', 'cdist', 'faiss'):
            raise NotImplementedError(f"Method {st} is not implemented")
        di = self.data_interim
        dev = self._b200_dev()
        n_raw = int(self.data_input_3d.idx_initial_src.shape[0])
        as_idx = lambda a: torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(dev, I64).contiguous()
        corres, labels = c2f.global_matches_from_3d(
            di.tile_pts_sub_feat_src, di.tile_pts_sub_feat_tgt, di.src_pts_sub, di.tgt_pts_sub,
            as_idx(di.idx_voxel2pts_src), as_idx(di.idx_voxel2pts_tgt), n_raw, float(self.para.max_magnitude))
        di.corres_3d_voxel_from_3d_idx = corres                    # :2889
        di.labels_from_3d = labels
        if getattr(self, "write_interim_files", True):             # :2898-2920, visualisation only
            self._write_global_3d_visualisation(labels)
        if getattr(self, "logging", None) is not None:
            self.logging.info('Global point matches from 3d is done!')

    def _write_global_3d_visualisation(self, labels):
        di = self.data_interim
        src = _dev_f32(di.src_pts_sub)
        tgt = _dev_f32(di.tgt_pts_sub, src.device)[labels.long()]
        mag = torch.linalg.norm(src - tgt, dim=1)
        keep = mag <= self.para.max_magnitude
        out = torch.hstack((src[keep], mag[keep, None]))
        if out.shape[0] > 1:
            out[0, 3] = 0
            out[1, 3] = {"rockfall_simulator": 0.06, "brienz_tls": 5, "mattertal": 10}.get(self.data.dataset, 10)
        os.makedirs(osp.join(self.output_root, 'results'), exist_ok=True)
        name = (f'c2f_dvfms_from_global_3d_src2tgt_wo_pruning_visualize_tile_{self.config.tile_id}.txt'
                if self.data.multiple_case else 'c2f_dvfms_from_global_3d_src2tgt_wo_pruning_visualize.txt')
        np.savetxt(osp.join(self.output_root, 'results', name), out.cpu())

    # ---- superpoint features (8f-2) ---------------------------------------------------------------------------
    def _compute_spt_feat_and_coord_with_fused_feats(self):
        """base.py:2561-2656: per-superpoint feature (self-attention pooling of the voxel descriptors + MLP) and
        centroid, for every superpoint of both epochs in one pass (nets.ClusterFeatureNetWithAttention)."""
        from . import nets
        di = self.data_interim
        dev = self._b200_dev()
        model = getattr(self, "feat_aggregate_model", None)
        if model is None or not hasattr(model, "aggregate_segments"):
            model = nets.ClusterFeatureNetWithAttention(input_feat_dim=64, hidden_feat_dim=64, output_feat_dim=64).to(dev)
            weight = osp.join(self.config.path_name.get("project_dir", ""), self.config.path_name.weight_dir,
                              self.config.path_name.pretrained_feature_aggregation_weight)
            state = torch.load(weight, map_location=dev)
            model.load_state_dict(state['state_dict'] if 'state_dict' in state else state)
            model.eval()
            self.feat_aggregate_model = model
        t = self._level_tables()
        with torch.no_grad():
            for name, ptr, idx in (("src", t.ptr_s, t.idx_s), ("tgt", t.ptr_t, t.idx_t)):
                p2v = di["idx_pts2voxel_" + name].to(dev)
                vox = p2v[idx.long()]                              # cluster_feature_net_self_attention.py:80-81
                ok = vox >= 0
                cnt = torch.zeros(ptr.numel() - 1, dtype=I64, device=dev)
                seg = torch.repeat_interleave(torch.arange(ptr.numel() - 1, device=dev), (ptr[1:] - ptr[:-1]).long(),
                                              output_size=int(idx.numel()))
                cnt.index_add_(0, seg[ok], torch.ones_like(seg[ok]))
                vptr = torch.zeros(ptr.numel(), dtype=I64, device=dev)
                vptr[1:] = torch.cumsum(cnt, 0)
                vsel = vox[ok]
                feats = _dev_f32(di["tile_pts_sub_feat_" + name], dev)[vsel].contiguous()
                coords = _dev_f32(di[name + "_pts_sub"], dev)[vsel].contiguous()
                f, c = model.aggregate_segments(feats, coords, vptr.to(I32))
                di["spt_feat_" + name], di["spt_coord_" + name] = f, c
        self._info(f'Computing superpoint coordinates and features is done! '
                   f'Num. of spt: {di.spt_feat_src.shape[0]} and {di.spt_feat_tgt.shape[0]}')

    # ---- B3 / B4 ----------------------------------------------------------------------------------------------
    def coarse_matching_with_different_types(self):
        """base.py:2925-3157.  3D: mutual feature-space NN under the coordinate gate; 2D: per source patch the
        target patch most of its 2D-lifted matches vote for; fusion: 2D pairs followed by 3D pairs."""
        self._info('Start coarse matching...')
        m_ = self.method
        di = self.data_interim
        t = self._level_tables()
        parts = {}
        if m_.coarse_matching_fusion or m_.coarse_matching_only_3d:
            if m_.feat_aggregate_type != 'learning_based':
                raise NotImplementedError
            if m_.get("use_img_patch_enhanced_3d_aggregation", False) or m_.get("use_img_pixel_enhanced_3d_aggregation", False):
                raise NotImplementedError("image-enhanced aggregation stays in the reference (image networks)")
            if not m_.get("use_normal_3d_aggregation", True):
                raise NotImplementedError
            self._compute_spt_feat_and_coord_with_fused_feats()
            if m_.coarse_refinement_3d_type not in ('only_max_mag', 'nn_mutual'):
                raise NotImplementedError
            parts["3d"] = c2f.coarse_matching_3d(di.spt_coord_src, di.spt_feat_src, di.spt_coord_tgt, di.spt_feat_tgt,
                                                 float(self.para.max_magnitude), m_.coarse_refinement_3d_type)
            if getattr(self, "logging", None) is not None:
                self.logging.info('Coarse matching from 3D source is done!')
        if m_.coarse_matching_fusion or m_.coarse_matching_only_2d:
            m2, j2, _ = c2f.coarse_matching_2d(di.corres_3d_from_2d_idx.to(t.idx_s.device), t.idx_s, t.ptr_s,
                                               di.idx_pts2spt_tgt, t.lab_t)
            parts["2d"] = (m2, j2)
            if getattr(self, "logging", None) is not None:
                self.logging.info('Coarse matching from 2D source is done!')
        if m_.coarse_matching_only_2d and m_.fine_matching_only_2d:
            m3, j3, _ = c2f.coarse_matching_2d(di.corres_3d_voxel_from_3d_idx.to(t.idx_s.device), t.idx_s, t.ptr_s,
                                               di.idx_pts2spt_tgt, t.lab_t)                   # :3072-3121 "extra 3d"
            parts["3d_extra"] = (m3, j3)
        self.spt_length = []                                       # :3124-3149
        if m_.coarse_matching_only_3d:
            order = ["3d"]
        elif m_.coarse_matching_only_2d:
            order = ["2d"]
        elif m_.coarse_matching_fusion:
            order = ["2d", "3d"]
        else:
            raise NotImplementedError
        if m_.coarse_matching_only_2d and m_.fine_matching_only_2d:
            order.append("3d_extra")
        for k in order:
            self.spt_length.append(int(parts[k][0].numel()))
        m = torch.cat([parts[k][0] for k in order])
        j = torch.cat([parts[k][1] for k in order])
        src_list = _IndexedList(t.list_s, m)
        tgt_list = _IndexedList(t.list_t, j)
        self.data_output.spt_corres_src = src_list                 # :3156-3157
        self.data_output.spt_corres_tgt = tgt_list
        _, pairs = self._b200_tables()
        pairs[_list_key(src_list)] = (m, j, t)

    # ---- F2 F3 D2 E1 D5 A4 ------------------------------------------------------------------------------------
    def _fine_config(self):
        m_ = self.method
        if m_.fine_matching_only_3d:
            mode = "only_3d"
        elif m_.fine_matching_only_2d:
            mode = "only_2d"
        elif m_.fine_matching_fusion:
            mode = "fusion"
        else:
            raise NotImplementedError
        if m_.get("weighting_svd", False):
            raise NotImplementedError("weighting_svd=True: the reference overwrites and then discards the weight vector "
                                      "(base.py:3290-3294,3326); no shipped config enables it")
        return pipeline.FineConfig(
            mode=mode, remove_low_quality_patch_matches=bool(m_.remove_low_quality_patch_matches),
            num_min_matches_for_quality_check=int(m_.num_min_matches_for_quality_check),
            thres_dist_diff=float(m_.thres_dist_diff), thres_inlier_ratio=float(m_.thres_inlier_ratio),
            num_min_fine_match=int(m_.num_min_fine_match), icp_refine=bool(m_.icp_refine), assign_type=m_.assign_type,
            output_tgt2src=bool(m_.output_tgt2src), icp_threshold=float(self.para.icp_threshold))

    def fine_matching_with_different_types(self):
        """base.py:3236-3457 as one launch sequence over all patch pairs of the level."""
        self._info('Start fine matching...')
        dev = self._b200_dev()
        do, di, d3 = self.data_output, self.data_interim, self.data_input_3d
        _, pairs = self._b200_tables()
        ent = pairs.get(_list_key(do.spt_corres_src)) if len(do.spt_corres_src) else None
        src, tgt = _dev_f32(d3.src_pts, dev), _dev_f32(d3.tgt_pts, dev)
        if ent is not None and int(ent[0].numel()) == len(do.spt_corres_src):
            m, j, t = ent
            sp_ptr, sp_idx, n_s = ops.gather_pairs_csr(t.ptr_s, t.idx_s, m)
            tp_ptr, tp_idx, n_t = ops.gather_pairs_csr(t.ptr_t, t.idx_t, j)
            tpo, pair_tgt = t.pop_t, j.to(I32).contiguous()
        else:                                                      # pair lists produced elsewhere: pack them
            def pack(lst):
                cnt = torch.tensor([0] + [int(x.numel()) for x in lst], dtype=I64)
                p = torch.cumsum(cnt, 0).to(dev, I32)
                idx = (torch.cat([x.reshape(-1) for x in lst]) if lst else torch.empty(0, dtype=I64)).to(dev, I32)
                return p, idx.contiguous(), int(idx.numel())
            sp_ptr, sp_idx, n_s = pack(do.spt_corres_src)
            tp_ptr, tp_idx, n_t = pack(do.spt_corres_tgt)
            # distinct target patches (a target patch matched by several source patches is one id)
            keys = {}
            pair_ids = []
            tpo = torch.full((tgt.shape[0],), -1, dtype=I32, device=dev)
            for x in do.spt_corres_tgt:
                k = (int(x[0]), int(x.numel())) if x.numel() else (-1, 0)
                if k not in keys:
                    keys[k] = len(keys)
                    tpo[x.to(dev).long()] = keys[k]
                pair_ids.append(keys[k])
            pair_tgt = torch.tensor(pair_ids, dtype=I32, device=dev)
        cfg = self._fine_config()
        need3d = cfg.mode in ("only_3d", "fusion")
        need2d = cfg.mode in ("only_2d", "fusion")
        med = float(getattr(self.para, "median_max_resolution", 0.0))
        r = ops.fine_matching(src, tgt, sp_idx, sp_ptr, tp_idx, tp_ptr, tpo, pair_tgt,
                              corr3d=di.corres_3d_voxel_from_3d_idx.to(dev).contiguous() if need3d else None,
                              corr2d=di.corres_3d_from_2d_idx.to(dev).contiguous() if need2d else None,
                              median_max_resolution=med, n_src_items=n_s, n_tgt_items=n_t, **cfg.fine_kwargs())
        self.fine_result = r
        dense, sparse, t2s = r.rows()
        if dense.shape[0] > 0 and cfg.icp_refine:                  # :3439-3451
            do.corres_3d_refine_apply_icp = dense
            if cfg.output_tgt2src:
                do.corres_3d_refine_apply_icp_tgt2src = t2s
            do.corres_3d_refine_apply_icp_discrete = sparse
        elif "corres_3d_refine_apply_icp" not in do:
            do.corres_3d_refine_apply_icp = []
        if isinstance(self.method.level_of_superpoint, list) and self.method.partition_type == 'superpoint':
            self._info(f'Fine matching is done for the superpoint level from {self.method.level_of_superpoint}!')
        else:
            self._info('Fine matching is done!')

    # ---- the entry point ----------------------------------------------------------------------------------------
    def implement_c2f_matching(self):
        """coarse_to_fine_matching.py:201-290, same stage order and the same fields."""
        m_ = self.method
        do, di = self.data_output, self.data_interim
        multi = m_.partition_type == 'superpoint' and isinstance(m_.level_of_superpoint, list)
        if m_.use_2d_matches and not (m_.coarse_matching_only_3d and m_.fine_matching_only_3d):
            self.global_matches_from_2d_with_different_types()     # image side: the reference's own method
        else:
            self._info('Skip 2d matching!')
        self._voxel_subsampling()
        self.implement_partition()
        self.load_partition()
        if multi:
            di.idx_spt2pts_src_multiple, di.idx_spt2pts_tgt_multiple = [], []
            spt_ids = []
            for level_current in m_.level_of_superpoint:
                di.idx_pts2spt_src = di.idx_pts2spt_src_multiple[level_current - 1]
                di.idx_pts2spt_tgt = di.idx_pts2spt_tgt_multiple[level_current - 1]
                self.prepare_pts2spt_dict()
                di.idx_spt2pts_src_multiple.append(di.idx_spt2pts_src)
                di.idx_spt2pts_tgt_multiple.append(di.idx_spt2pts_tgt)
                spt_ids.append((di.idx_spt_src, di.idx_spt_tgt))
        else:
            self.prepare_pts2spt_dict()
        if getattr(self, "debugging", None) is not None and self.debugging.get("use_debugging", False):
            self.start_debugging('reduce_num_spt')
        if m_.coarse_matching_only_2d and m_.fine_matching_only_2d:
            self._info('Skip computing point features!')
        else:
            self.compute_point_feat()
            self.global_matches_from_3d()
        if multi:
            do.spt_corres_src_multiple, do.spt_corres_tgt_multiple = [], []
            for level_current in m_.level_of_superpoint:
                di.idx_pts2spt_src = di.idx_pts2spt_src_multiple[level_current - 1]
                di.idx_pts2spt_tgt = di.idx_pts2spt_tgt_multiple[level_current - 1]
                di.idx_spt2pts_src = di.idx_spt2pts_src_multiple[level_current - 1]
                di.idx_spt2pts_tgt = di.idx_spt2pts_tgt_multiple[level_current - 1]
                di.idx_spt_src, di.idx_spt_tgt = spt_ids[level_current - 1]    # (the reference leaves the last level's ids here)
                self.coarse_matching_with_different_types()
                do.spt_corres_src_multiple.append(do.spt_corres_src)
                do.spt_corres_tgt_multiple.append(do.spt_corres_tgt)
        else:
            self.coarse_matching_with_different_types()
        if multi:
            do.corres_3d_refine_apply_icp_multiple = []
            if m_.output_tgt2src:
                do.corres_3d_refine_apply_icp_tgt2src_multiple = []
            do.corres_3d_refine_apply_icp_discrete_multiple = []
            self.fine_results_multiple = []                        # per-level FineResult (transforms, status, iterations)
            for level_current in m_.level_of_superpoint:
                do.spt_corres_src = do.spt_corres_src_multiple[level_current - 1]
                do.spt_corres_tgt = do.spt_corres_tgt_multiple[level_current - 1]
                self.fine_matching_with_different_types()
                self.fine_results_multiple.append(self.fine_result)
                do.corres_3d_refine_apply_icp_multiple.append(do.corres_3d_refine_apply_icp)
                if m_.output_tgt2src:
                    do.corres_3d_refine_apply_icp_tgt2src_multiple.append(do.corres_3d_refine_apply_icp_tgt2src)
                do.corres_3d_refine_apply_icp_discrete_multiple.append(do.corres_3d_refine_apply_icp_discrete)
        else:
            self.fine_matching_with_different_types()
        if multi:
            self._info('Start merging correspondences...')
            merge = c2f.merge_correspondences_by_priority_with_distance_threshold
            do.corres_3d_refine_apply_icp = merge(do.corres_3d_refine_apply_icp_multiple)
            if m_.output_tgt2src:
                do.corres_3d_refine_apply_icp_tgt2src = merge(do.corres_3d_refine_apply_icp_tgt2src_multiple)
            do.corres_3d_refine_apply_icp_discrete = merge(do.corres_3d_refine_apply_icp_discrete_multiple)
        if not (isinstance(do.corres_3d_refine_apply_icp, list) and do.corres_3d_refine_apply_icp == []):
            self.save_process_dvf()


class StandaloneBase:
    """The parts of Coarse2Fine_Base around the hot methods, for use WITHOUT the reference tree (tests, benchmarks):
    config plumbing, tile / partition / feature readers, result writers.  Same file names and formats.
    Image matching and partitioning are the reference's business (north_star: "stay as in the reference")."""

    def __init__(self, config):
        self.config = config
        self.logging = config.get("logging", None)
        self.verbose = config.get("verbose", False)
        self.save_interim = config.get("save_interim", False)
        self.device = config.get("device", "cuda")
        self.debugging = config.get("debugging", edict(use_debugging=False))
        self.write_interim_files = config.get("write_interim_files", True)
        # in-memory tile (tests, benchmarks, callers that already hold the tile): config.tile_tensors = dict with
        #   src_pts, tgt_pts (n,3); partition_src, partition_tgt: (n,) labels or a list of them (one per level);
        #   feat_src, feat_tgt (n_sub,D) descriptors of the sub-sampled clouds, OR feat_raw_src, feat_raw_tgt (n,D)
        #   per raw point (gathered through idx_voxel2pts_*); optionally corres_3d_from_2d_idx (n,2) int64
        self.tile_tensors = config.get("tile_tensors", None)
        if config.get("feat_aggregate_model", None) is not None:
            self.feat_aggregate_model = config.feat_aggregate_model
        self._initialize()
        self._read_data()

    def _initialize(self):                                         # base.py:644-659
        self.data_input_2d, self.data_input_3d, self.data_interim, self.data_output = edict(), edict(), edict(), edict()
        self.backbones = edict()
        self.input_root = self.config.path_name.input_root
        self.output_root = self.config.path_name.output_root
        self.data = self.config.data
        self.method = self.config.method
        self.para = self.config.parameter_setting
        self.visualize = self.config.get("visualization", edict())

    def _read_data(self):                                          # base.py:890-916 (3D part)
        from .piecewise_icp import _read_xyz
        dev = torch.device(self.device) if not isinstance(self.device, torch.device) else self.device
        if self.tile_tensors is not None:
            self.src_pcd_path = self.tgt_pcd_path = None
        elif self.data.multiple_case:
            self.src_pcd_path, self.tgt_pcd_path = self.config.src_tile_overlap_path, self.config.tgt_tile_overlap_path
        else:
            self.src_pcd_path = osp.join(self.input_root, 'raw_pcd', self.data.src_pcd)
            self.tgt_pcd_path = osp.join(self.input_root, 'raw_pcd', self.data.tgt_pcd)
        d3 = self.data_input_3d
        tt = self.tile_tensors
        for name, path in (("src", self.src_pcd_path), ("tgt", self.tgt_pcd_path)):
            if tt is not None:
                p = tt[name + "_pts"]
                p = p if torch.is_tensor(p) else torch.from_numpy(np.asarray(p))
                p = p.to(dev, non_blocking=True)
                d3[name + "_pcd"] = _Cloud(p)
                d3[name + "_pts"] = p.float().contiguous()
                continue
            pts64 = _read_xyz(path)
            d3[name + "_pcd"] = _Cloud(pts64)
            d3[name + "_pts"] = torch.from_numpy(pts64).float().to(dev)       # pcd2tensor: float32
        d3.idx_initial_src = torch.arange(d3.src_pts.shape[0], device=dev)
        d3.idx_initial_tgt = torch.arange(d3.tgt_pts.shape[0], device=dev)
        if not (self.method.coarse_matching_only_3d and self.method.fine_matching_only_3d):
            c = self.config.get("corres_3d_from_2d_idx", None)
            if c is None and tt is not None:
                c = tt.get("corres_3d_from_2d_idx", None)
            if c is None:
                raise NotImplementedError("2D matching / 2D->3D lifting run in the reference (image networks); hand the "
                                          "lifted matches in as config.corres_3d_from_2d_idx (N,2) int64")
            self.data_interim.corres_3d_from_2d_idx = torch.as_tensor(c).to(dev, I64)

    def global_matches_from_2d_with_different_types(self):
        if "corres_3d_from_2d_idx" not in self.data_interim:
            raise NotImplementedError("image matching stays in the reference")

    def implement_partition(self):                                 # base.py:2658-2714
        if self.method.partition:
            raise NotImplementedError("supervoxel / superpoint partitioning stays in the reference; "
                                      "set method.partition: False to load its result files")
        if self.logging is not None:
            self.logging.info('Skip the partition process. The partition result will be loaded from path.')

    def load_partition(self):                                      # base.py:1237-1299
        m_ = self.method
        if self.tile_tensors is not None and "partition_src" in self.tile_tensors:
            dev = self.data_input_3d.src_pts.device
            di = self.data_interim
            get = lambda a: [torch.as_tensor(x).to(dev, I64) for x in a] if isinstance(a, (list, tuple)) else torch.as_tensor(a).to(dev, I64)
            ps, pt = get(self.tile_tensors["partition_src"]), get(self.tile_tensors["partition_tgt"])
            if m_.partition_type == 'superpoint' and isinstance(m_.level_of_superpoint, list):
                di.idx_pts2spt_src_multiple = [ps[lv - 1] for lv in m_.level_of_superpoint] if isinstance(ps, list) else [ps]
                di.idx_pts2spt_tgt_multiple = [pt[lv - 1] for lv in m_.level_of_superpoint] if isinstance(pt, list) else [pt]
            elif m_.partition_type == 'superpoint' and isinstance(ps, list):
                di.idx_pts2spt_src, di.idx_pts2spt_tgt = ps[m_.level_of_superpoint - 1], pt[m_.level_of_superpoint - 1]
            else:
                di.idx_pts2spt_src, di.idx_pts2spt_tgt = ps, pt
            return
        folder = f'{m_.partition_type}_partition'
        pdir = osp.join(self.output_root, folder)
        if not os.path.isdir(pdir) or not os.listdir(pdir):
            raise FileNotFoundError(f"No partition result in '{pdir}'")
        suffix = f'_tile_{self.config.tile_id}' if self.data.multiple_case else ''
        ps = np.loadtxt(osp.join(pdir, f'partition_of_input_src{suffix}.txt'))
        pt = np.loadtxt(osp.join(pdir, f'partition_of_input_tgt{suffix}.txt'))
        dev = self.data_input_3d.src_pts.device
        col = lambda a, c: torch.from_numpy(a[:, c]).to(I64).to(dev)
        di = self.data_interim
        if m_.partition_type == 'superpoint' and isinstance(m_.level_of_superpoint, int):
            di.idx_pts2spt_src, di.idx_pts2spt_tgt = col(ps, 2 + 4 * m_.level_of_superpoint), col(pt, 2 + 4 * m_.level_of_superpoint)
        elif m_.partition_type == 'superpoint' and isinstance(m_.level_of_superpoint, list):
            di.idx_pts2spt_src_multiple = [col(ps, 2 + 4 * lv) for lv in m_.level_of_superpoint]
            di.idx_pts2spt_tgt_multiple = [col(pt, 2 + 4 * lv) for lv in m_.level_of_superpoint]
        else:
            di.idx_pts2spt_src, di.idx_pts2spt_tgt = col(ps, 6), col(pt, 6)

    @property
    def _feat_path(self):                                          # base.py:998-1004
        if self.data.multiple_case:
            return osp.join(self.output_root, 'features', f'features_tile_{self.config.tile_id}.npz')
        return osp.join(self.output_root, 'features', 'features.npz')

    def compute_point_feat(self):                                  # base.py:1965-2072
        di = self.data_interim
        dev = self.data_input_3d.src_pts.device
        tt = self.tile_tensors
        if tt is not None and ("feat_src" in tt or "feat_raw_src" in tt):
            if "feat_raw_src" in tt:                               # stands in for the descriptor network: like :1982 the
                self._compute_median_resolution()                  # patch radius is taken from the sub-sampled clouds
            for name in ("src", "tgt"):
                if "feat_" + name in tt:
                    f = torch.as_tensor(tt["feat_" + name]).to(dev)
                else:                                              # descriptor of a voxel = descriptor of its raw point
                    f = torch.as_tensor(tt["feat_raw_" + name]).to(dev)[torch.as_tensor(di["idx_voxel2pts_" + name]).to(dev).long()]
                di["tile_pts_sub_feat_" + name] = f.float().contiguous()
            return
        if not self.method.point_feat_compute:
            if not osp.exists(self._feat_path):
                raise FileNotFoundError(f"The feature path '{self._feat_path}' is not found")
            f = np.load(self._feat_path)
            di.tile_pts_sub_feat_src = torch.from_numpy(f['src_feat']).to(dev)
            di.tile_pts_sub_feat_tgt = torch.from_numpy(f['tgt_feat']).to(dev)
            return
        if self.method.feat_type != 'DIPs':
            raise NotImplementedError
        from .data_loader import Preprocess_Dataset
        radius = np.sqrt(3) * (10 * self._compute_median_resolution())         # :1982
        net = self.config.feat_desc_nn
        out = {}
        for name in ("src", "tgt"):
            ds = Preprocess_Dataset(di[name + "_pcd_sub"], di[name + "_pcd_sub"], self.para.points_per_batch, radius)
            feats = []
            for b in range(len(ds)):
                x = ds[b]
                x = x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x))
                feats.append(net(x.to(dev).float())[0])
            out[name] = torch.cat(feats, dim=0)
        di.tile_pts_sub_feat_src, di.tile_pts_sub_feat_tgt = out["src"], out["tgt"]

    def save_process_dvf(self):                                    # base.py:3459-3655 (src2tgt outputs)
        if not self.method.icp_refine:
            return
        do = self.data_output
        mag = lambda rows: torch.linalg.norm(rows[:, 3:6] - rows[:, :3], dim=1)[:, None]
        do.corres_3d_magnitude_refine_apply_icp = mag(do.corres_3d_refine_apply_icp)
        if self.method.output_tgt2src:
            do.corres_3d_magnitude_refine_apply_icp_tgt2src = mag(do.corres_3d_refine_apply_icp_tgt2src)
        do.corres_3d_magnitude_refine_apply_icp_discrete = mag(do.corres_3d_refine_apply_icp_discrete)
        if not self.config.get("write_results", True):
            return
        res = osp.join(self.output_root, 'results')
        os.makedirs(res, exist_ok=True)
        vmax = {"rockfall_simulator": 0.06, "brienz_tls": 5, "mattertal": 10}.get(self.data.dataset, 10)
        dense = do.corres_3d_refine_apply_icp.cpu().numpy()
        m = do.corres_3d_magnitude_refine_apply_icp.cpu().numpy()
        if self.data.multiple_case:
            tid = self.config.tile_id
            names = (f'c2f_dense_dvfs_src2tgt_tile_{tid}.txt', f'c2f_dense_dvfms_src2tgt_tile_{tid}.txt',
                     f'c2f_dense_dvfms_src2tgt_visualize_tile_{tid}.txt', f'c2f_sparse_dvfms_src2tgt_visualize_tile_{tid}.txt')
        else:
            names = ('c2f_dvfs_src2tgt.txt', 'c2f_dvfms_src2tgt.txt', 'c2f_dvfms_src2tgt_visualize_0_5.txt',
                     'c2f_dvfms_src2tgt_discrete_visualize_0_5.txt')
        np.savetxt(osp.join(res, names[0]), dense, delimiter=' ', fmt='%.6f')
        np.savetxt(osp.join(res, names[1]), np.hstack((dense[:, :3], m)), delimiter=' ', fmt='%.6f')
        vis = m.copy()
        if vis.shape[0] > 1:
            vis[0], vis[1] = 0, vmax                                # quirk q10
        np.savetxt(osp.join(res, names[2]), np.hstack((dense[:, :3], vis)), delimiter=' ', fmt='%.6f')
        sp = do.corres_3d_refine_apply_icp_discrete.cpu().numpy()
        ms = do.corres_3d_magnitude_refine_apply_icp_discrete.cpu().numpy()
        if ms.shape[0] > 1:
            ms[0], ms[1] = 0, vmax
        np.savetxt(osp.join(res, names[3]), np.hstack((sp[:, :3], ms)), delimiter=' ', fmt='%.6f')


def bind(base):
    """class Coarse2Fine(HotPathMixin, base): the hot methods replace `base`'s, everything else is inherited."""
    class Coarse2Fine(HotPathMixin, base):
        def __init__(self, config):
            super().__init__(config)
            if not hasattr(self, "debugging"):
                self.debugging = config.get("debugging", edict(use_debugging=False)) if hasattr(config, "get") else config.debugging
    Coarse2Fine.__doc__ = "Drop-in for `from src.coarse_to_fine_matching import Coarse2Fine` (main_fusion.py:8,147-148)."
    return Coarse2Fine


Coarse2Fine = bind(StandaloneBase)
