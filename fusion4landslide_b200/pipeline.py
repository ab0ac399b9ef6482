"""Tile-level public API of the hot path: device-resident inputs -> device-resident DVF.

`prepare_tile` is the step immediately BEFORE the path (labels -> CSR, the reference's
`prepare_pts2spt_dict`, base.py:1301-1351) and is not timed as part of it; `displacement_field`
is the path: A1 median resolution (k=2 self-kNN on both epochs) + the fused fine-matching stage.
`displacement_field_host` is the same call for a caller that holds HOST buffers (pinned):
host->device copies of the inputs and device->host copies of the results are part of the call.
"""
import os

import torch

from . import ops, synth


class FineConfig:
    """The hot-path keys of configs/landslide/fusion_3d_brienz.yaml (method.* / parameter_setting.*)."""

    def __init__(self, mode="only_3d", remove_low_quality_patch_matches=True,
                 num_min_matches_for_quality_check=10, thres_dist_diff=0.5, thres_inlier_ratio=0.15,
                 num_min_fine_match=10, icp_refine=True, assign_type="assign_then_nn",
                 output_tgt2src=False, icp_threshold=0.1, icp_max_iter=30,
                 num_min_matches_for_small_patch=10):
        self.__dict__.update({k: v for k, v in locals().items() if k != "self"})

    def fine_kwargs(self):
        d = dict(self.__dict__)
        d.pop("num_min_matches_for_small_patch")
        return d


class TileInputs:
    """Per-tile device tensors the path consumes (SURVEY 9.1 names in comments)."""
    __slots__ = ("src", "tgt", "corr3d", "corr2d", "sp_idx", "sp_ptr", "tp_idx", "tp_ptr",
                 "tgt_patch_of_point", "pair_tgt_patch", "n_src_items", "n_tgt_items", "n_pairs")

    def tensors(self):
        return [(k, getattr(self, k)) for k in self.__slots__ if torch.is_tensor(getattr(self, k))]

    def nbytes(self):
        return sum(t.numel() * t.element_size() for _, t in self.tensors())


def prepare_tile(src, tgt, label_src, label_tgt, corr3d, corr2d=None, min_pts=10, pairs=None):
    """labels -> CSR patch lists + matched patch pairs, on the GPU (f4l_labels_to_csr, f4l_gather_pairs_csr).
    `pairs` = (src patch pos, tgt patch pos) from the coarse matching; default: patches carrying the same
    label (the synthetic ground-truth pairing).  CPU tensors are handled by synth.prepare_tile_host (the CPU
    arm of bench.py builds the oracle's inputs with it)."""
    if not src.is_cuda:
        return synth.prepare_tile_host(src, tgt, label_src, label_tgt, corr3d, corr2d, min_pts, pairs, TileInputs)
    dev = src.device
    lab_s, ptr_s, idx_s, _ = ops.labels_to_csr(label_src.to(dev, torch.int64).contiguous(), min_pts)   # idx_spt2pts_src
    lab_t, ptr_t, idx_t, tpo = ops.labels_to_csr(label_tgt.to(dev, torch.int64).contiguous(), min_pts)  # idx_spt2pts_tgt
    m, j = pairs if pairs is not None else synth.pair_patches(lab_s, lab_t)
    t = TileInputs()
    t.src, t.tgt, t.corr3d, t.corr2d = src.contiguous(), tgt.contiguous(), corr3d.contiguous(), corr2d
    t.sp_ptr, t.sp_idx, t.n_src_items = ops.gather_pairs_csr(ptr_s, idx_s, m)
    t.tp_ptr, t.tp_idx, t.n_tgt_items = ops.gather_pairs_csr(ptr_t, idx_t, j)
    t.tgt_patch_of_point = tpo.contiguous()
    t.pair_tgt_patch = j.to(torch.int32).contiguous()
    t.n_pairs = int(m.numel())
    return t


def displacement_field(tile, cfg=None, out=None, med_out=None, peer_dense=None, side_stream=None):
    """The hot path on one tile, device -> device, no host synchronisation.
    Returns (FineResult, median_resolution device scalar).  peer_dense: see ops.fine_matching.
    side_stream: run A1 there, concurrently with correspondence selection and the rigid fits of the same tile
    (only the assign step needs the resolution; the library waits for it in-stream)."""
    cfg = cfg or FineConfig()
    ev = None
    if side_stream is None:
        med = ops.median_resolution(tile.src, tile.tgt, out=med_out)                      # A1
    else:
        cur = torch.cuda.current_stream(tile.src.device)
        if med_out is None:
            med_out = torch.empty((1,), dtype=torch.float32, device=tile.src.device)
        side_stream.wait_stream(cur)
        with torch.cuda.stream(side_stream):
            med = ops.median_resolution(tile.src, tile.tgt, out=med_out)
            ev = torch.cuda.Event()
            ev.record(side_stream)
    r = ops.fine_matching(tile.src, tile.tgt, tile.sp_idx, tile.sp_ptr, tile.tp_idx, tile.tp_ptr,
                          tile.tgt_patch_of_point, tile.pair_tgt_patch, corr3d=tile.corr3d,
                          corr2d=tile.corr2d, d_median_resolution=med, n_src_items=tile.n_src_items,
                          n_tgt_items=tile.n_tgt_items, out=out, peer_dense=peer_dense, median_event=ev,
                          **cfg.fine_kwargs())
    return r, med


def make_streams(n, device):
    """Side streams for tile-level concurrency (tiles are independent; the library keeps one workspace per
    stream)."""
    return [torch.cuda.Stream(device=device) for _ in range(max(1, int(n)))]


def displacement_field_tiles(tiles, cfg=None, outs=None, meds=None, streams=None, peers=None, side_streams=None,
                             push=None):
    """The hot path over many tiles.  With `streams`, tile i runs on streams[i % n]: the small kernels and the
    tail of one tile's patch loop overlap with the next tile's work.  The caller's current stream waits for all
    of them at the end, so events recorded around this call time the whole batch.  `peers[i]`: peer pointers
    of tile i's dense slot (exchange.PeerExchange.peer_ptrs) -- the dense rows then land in every GPU's
    gathered field while the tile is computed."""
    cfg = cfg or FineConfig()
    res = []

    def push_rows(i, r, st):
        # push = (exchange stream, per-tile pointer lists): the finished dense rows of tile i go to every peer's field
        # through the copy kernel (ops.peer_push) on the exchange stream, while the next tiles are computed
        xs, ptr_lists = push
        ev = torch.cuda.Event()
        ev.record(st)
        with torch.cuda.stream(xs):
            xs.wait_event(ev)
            ops.peer_push(r.dense, r.counts, ptr_lists[i])

    if not streams or len(streams) == 1 and streams[0] is None:
        cur = torch.cuda.current_stream(tiles[0].src.device) if tiles else None
        if push is not None and tiles:
            push[0].wait_stream(cur)
        for i, t in enumerate(tiles):
            res.append(displacement_field(t, cfg, out=None if outs is None else outs[i],
                                          med_out=None if meds is None else meds[i:i + 1],
                                          peer_dense=None if peers is None else peers[i]))
            if push is not None:
                push_rows(i, res[-1][0], cur)
        if push is not None and tiles:
            cur.wait_stream(push[0])
        return res
    cur = torch.cuda.current_stream(tiles[0].src.device) if tiles else None
    for s in streams:
        s.wait_stream(cur)
    if push is not None:
        push[0].wait_stream(cur)
    if os.environ.get("F4L_TILE_ORDER", "") == "phased":
        # experiment: every tile's A1 chain first, then every tile's fine matching (same stream per tile)
        med_l = []
        for i, t in enumerate(tiles):
            with torch.cuda.stream(streams[i % len(streams)]):
                med_l.append(ops.median_resolution(t.src, t.tgt, out=None if meds is None else meds[i:i + 1]))
        for i, t in enumerate(tiles):
            with torch.cuda.stream(streams[i % len(streams)]):
                r = ops.fine_matching(t.src, t.tgt, t.sp_idx, t.sp_ptr, t.tp_idx, t.tp_ptr, t.tgt_patch_of_point,
                                      t.pair_tgt_patch, corr3d=t.corr3d, corr2d=t.corr2d, d_median_resolution=med_l[i],
                                      n_src_items=t.n_src_items, n_tgt_items=t.n_tgt_items,
                                      out=None if outs is None else outs[i],
                                      peer_dense=None if peers is None else peers[i], **cfg.fine_kwargs())
                res.append((r, med_l[i]))
    else:
        for i, t in enumerate(tiles):
            with torch.cuda.stream(streams[i % len(streams)]):
                res.append(displacement_field(t, cfg, out=None if outs is None else outs[i],
                                              med_out=None if meds is None else meds[i:i + 1],
                                              peer_dense=None if peers is None else peers[i],
                                              side_stream=None if not side_streams else side_streams[i % len(side_streams)]))
                if push is not None:
                    push_rows(i, res[-1][0], streams[i % len(streams)])
    for s in streams:
        cur.wait_stream(s)
    if push is not None:
        cur.wait_stream(push[0])
    return res


def displacement_field_tiles_batched(tiles, cfg=None, outs=None, meds=None, streams=None, peers=None, side_streams=None,
                                     cache=None, groups=1, fit_stream=None, fit_ctas_per_sm=0):
    """The hot path over many tiles with the rigid fits of a GROUP of tiles in one persistent launch.

    Per tile (on its stream): A1 on a side stream, correspondence selection.  Per group of tiles, on `fit_stream`: one
    launch in which every resident warp draws the next patch pair of any tile of the group from a device-side queue
    (ops.fine_fit_tiles: no per-tile wave quantisation, one tail per group).  Then per tile again: the large pairs,
    apply + assign, sparse rows.  With groups > 1 the fit launch of one group runs while the tiles of the previous
    group finish and those of the next one select (the fit kernel is issue-bound, the others memory/latency-bound).
    Same results as displacement_field_tiles (the per-pair computation is the same device function).
    cache: a dict the caller keeps between steps with static buffers (benchmarks, CUDA-graph capture): the prepared
    calls, their workspaces and events are then built once."""
    cfg = cfg or FineConfig()
    if not tiles:
        return []
    dev = tiles[0].src.device
    cur = torch.cuda.current_stream(dev)
    streams = streams or [cur]
    n = len(streams)
    fit_stream = fit_stream or cur
    for s in list(streams) + [fit_stream]:
        if s is not cur:
            s.wait_stream(cur)
    groups = max(1, min(int(groups), len(tiles)))
    per = (len(tiles) + groups - 1) // groups
    res = [None] * len(tiles)
    group_idx = [list(range(g0, min(g0 + per, len(tiles)))) for g0 in range(0, len(tiles), per)]

    def select(idx):
        entries = []
        for i in idx:
            t = tiles[i]
            st = streams[i % n]
            ent = None if cache is None else cache.get(i)
            with torch.cuda.stream(st):
                if ent is None:
                    med = (meds[i:i + 1] if meds is not None else torch.empty((1,), dtype=torch.float32, device=dev))
                    side = side_streams[i % len(side_streams)] if side_streams else None
                    ev = torch.cuda.Event() if side is not None else None
                    call = ops.fine_prepare(t.src, t.tgt, t.sp_idx, t.sp_ptr, t.tp_idx, t.tp_ptr, t.tgt_patch_of_point,
                                            t.pair_tgt_patch, corr3d=t.corr3d, corr2d=t.corr2d, d_median_resolution=med,
                                            n_src_items=t.n_src_items, n_tgt_items=t.n_tgt_items,
                                            out=None if outs is None else outs[i],
                                            peer_dense=None if peers is None else peers[i], median_event=ev,
                                            own_workspace=True, **cfg.fine_kwargs())
                    ent = (call, med, side, ev)
                    if cache is not None:
                        cache[i] = ent
                call, med, side, ev = ent
                if side is None:
                    ops.median_resolution(t.src, t.tgt, out=med)                               # A1
                else:
                    side.wait_stream(st)
                    with torch.cuda.stream(side):
                        ops.median_resolution(t.src, t.tgt, out=med)
                        ev.record(side)
                call.run(ops.PHASE_SELECT)
                sel_ev = torch.cuda.Event()
                sel_ev.record(st)
            entries.append(ent + (sel_ev,))
        return entries

    # software pipeline over the groups: the selections of group g+1 are enqueued on the tile streams BEFORE the finish
    # phases of group g, so the fit launch of g+1 follows that of g on the fit stream while group g finishes
    pending = select(group_idx[0])
    for g, idx in enumerate(group_idx):
        entries = pending
        with torch.cuda.stream(fit_stream):
            for e in entries:
                fit_stream.wait_event(e[4])
            ops.fine_fit_tiles([e[0] for e in entries], fit_ctas_per_sm)
            fit_ev = torch.cuda.Event()
            fit_ev.record(fit_stream)
        if g + 1 < len(group_idx):
            pending = select(group_idx[g + 1])
        for i, (call, med, side, ev, sel_ev) in zip(idx, entries):
            st = streams[i % n]
            with torch.cuda.stream(st):
                st.wait_event(fit_ev)
                res[i] = (call.run(ops.PHASE_FIT_LARGE | ops.PHASE_FINISH), med)
    for s in list(streams) + [fit_stream] + list(side_streams or []):
        if s is not cur:
            cur.wait_stream(s)
    return res


class HostTile:
    """Pinned host copy of a TileInputs (what a caller holding numpy / CPU tensors passes)."""

    def __init__(self, tile):
        self.meta = (tile.n_src_items, tile.n_tgt_items, tile.n_pairs)
        self.t = {k: v.cpu().pin_memory() for k, v in tile.tensors()}

    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in self.t.values())


def displacement_field_host(host_tile, cfg=None, device="cuda:0", host_out=None):
    """Same call with HOST buffers: copies the inputs to the device, runs the path, copies the dense
    DVF rows, per-patch transforms / status and the row counts back.  Returns a dict of pinned host
    tensors (views sized by the true row counts) and the bytes moved (h2d, d2h)."""
    dev = torch.device(device)
    t = TileInputs()
    for k in TileInputs.__slots__:
        setattr(t, k, None)
    h2d = 0
    for k, v in host_tile.t.items():
        setattr(t, k, v.to(dev, non_blocking=True))
        h2d += v.numel() * v.element_size()
    t.n_src_items, t.n_tgt_items, t.n_pairs = host_tile.meta
    r, med = displacement_field(t, cfg)
    counts = r.counts.cpu()                      # sync: the row counts size the copies below
    nd, nsp = int(counts[0]), int(counts[1])
    if host_out is None:
        host_out = {}
    res = {}
    d2h = 16
    for name, ten in (("dense", r.dense[:nd]), ("sparse", r.sparse[:nsp]), ("T", r.T), ("status", r.status)):
        buf = host_out.get(name)
        if buf is None or buf.shape[0] < ten.shape[0]:
            buf = torch.empty(ten.shape, dtype=ten.dtype).pin_memory()
            host_out[name] = buf
        view = buf[:ten.shape[0]]
        view.copy_(ten, non_blocking=True)
        res[name] = view
        d2h += ten.numel() * ten.element_size()
    res["median_resolution"] = med.cpu()
    d2h += 4
    torch.cuda.current_stream(dev).synchronize()
    return res, h2d, d2h


class HostPipeline:
    """Host-buffer API over many tiles: pinned host inputs -> device, the path, results -> pinned host, with the
    copies of one tile overlapping the kernels of the others (one side stream per slot).  Row counts are read
    back per tile (a stream-local synchronisation) so that only the rows that exist cross PCIe."""

    def __init__(self, host_tiles, cfg=None, device="cuda:0", n_streams=4, want_sparse=True, sparse_once=False,
                 expand_threads=8, compact_corr=False, pack_threads=4):
        """sparse_once: the reference appends every pair's sparse rows twice (base.py:3430,3436); with this option
        the kernels emit them once and host threads restore the doubled layout while later tiles are in flight
        (same bytes in the returned tensors, a third less device->host traffic).  OFF by default: measured on the
        B200 box the DMA engines move the second copy at 56 GB/s while host threads rebuild it at ~14 GB/s
        (tools/exp_e2e.py: 25.2 ms vs 22.6 ms per 16 tiles) -- it only pays on hosts with a slow PCIe link."""
        # compact_corr: the fused stage reads only column 1 of the (n,2) int64 correspondence tables; host threads repack
        # it to int32 (n) into pinned buffers while earlier tiles are in flight, and 4 instead of 16 bytes per source
        # point cross PCIe (the conversion is part of run(), i.e. inside whatever the caller times)
        self.compact_corr = bool(compact_corr)
        self.pack_pool = None
        self.corr32 = []
        if self.compact_corr:
            import concurrent.futures
            self.pack_pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, int(pack_threads)))
            for ht in host_tiles:
                self.corr32.append({k: torch.empty((ht.t[k].shape[0],), dtype=torch.int32).pin_memory()
                                    for k in ("corr3d", "corr2d") if ht.t.get(k) is not None})
        self.host_tiles = host_tiles
        self.cfg = cfg or FineConfig()
        self.sparse_once = bool(sparse_once and want_sparse and self.cfg.fine_kwargs().get("assign_type") == "assign_then_nn")
        self.pool = None
        if self.sparse_once:
            import concurrent.futures
            self.pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, int(expand_threads)))
        self.dev = torch.device(device)
        self.streams = make_streams(n_streams, self.dev)
        self.want_sparse = want_sparse
        self.out = []
        for ht in host_tiles:
            n_s, _, q = ht.meta
            o = {"dense": torch.empty((n_s, 6), dtype=torch.float32).pin_memory(),
                 "T": torch.empty((q, 4, 4), dtype=torch.float32).pin_memory(),
                 "status": torch.empty((q,), dtype=torch.int8).pin_memory(),
                 "counts": torch.empty((4,), dtype=torch.int32).pin_memory(),
                 "median_resolution": torch.empty((1,), dtype=torch.float32).pin_memory()}
            if want_sparse:
                o["sparse"] = torch.empty((2 * n_s, 6), dtype=torch.float32).pin_memory()
            if self.sparse_once:
                o["sparse_once"] = torch.empty((n_s, 6), dtype=torch.float32).pin_memory()
                o["pair_rows"] = torch.empty((q,), dtype=torch.int32).pin_memory()
            self.out.append(o)

    def run(self):
        """Returns (list of per-tile dicts of host views, h2d bytes, d2h bytes)."""
        cur = torch.cuda.current_stream(self.dev)
        for s in self.streams:
            s.wait_stream(cur)
        h2d = d2h = 0
        pending = []
        self._expanding = []
        self._futures = []
        kw = dict(self.cfg.fine_kwargs())
        if self.sparse_once:
            kw["assign_type"] = "assign_then_nn_once"
        packs = []
        if self.compact_corr:                              # all repacks are queued now; tile i waits for its own only
            for i, ht in enumerate(self.host_tiles):
                packs.append({k: self.pack_pool.submit(ops.host_pack_corr_targets, ht.t[k], buf, 1)
                              for k, buf in self.corr32[i].items()})
        for i, ht in enumerate(self.host_tiles):
            st = self.streams[i % len(self.streams)]
            with torch.cuda.stream(st):
                t = TileInputs()
                for k in TileInputs.__slots__:
                    setattr(t, k, None)
                c32 = {}
                for k, v in ht.t.items():
                    if self.compact_corr and k in self.corr32[i]:
                        packs[i][k].result()
                        v = self.corr32[i][k]
                        c32[k] = v.to(self.dev, non_blocking=True)
                    else:
                        setattr(t, k, v.to(self.dev, non_blocking=True))
                    h2d += v.numel() * v.element_size()
                t.n_src_items, t.n_tgt_items, t.n_pairs = ht.meta
                med = ops.median_resolution(t.src, t.tgt)
                r = ops.fine_matching(t.src, t.tgt, t.sp_idx, t.sp_ptr, t.tp_idx, t.tp_ptr, t.tgt_patch_of_point,
                                      t.pair_tgt_patch, corr3d=t.corr3d, corr2d=t.corr2d, d_median_resolution=med,
                                      n_src_items=t.n_src_items, n_tgt_items=t.n_tgt_items,
                                      corr3d_tgt=c32.get("corr3d"), corr2d_tgt=c32.get("corr2d"), **kw)
                t.corr3d, t.corr2d = (c32.get("corr3d") if t.corr3d is None else t.corr3d,
                                      c32.get("corr2d") if t.corr2d is None else t.corr2d)   # kept alive with the tile
                o = self.out[i]
                o["counts"].copy_(r.counts, non_blocking=True)
                if self.sparse_once:
                    o["pair_rows"].copy_(r.sparse_pair_rows, non_blocking=True)
                o["median_resolution"].copy_(med, non_blocking=True)
                o["T"].copy_(r.T, non_blocking=True)
                o["status"].copy_(r.status, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st)
            pending.append((i, st, ev, r, t))
            # drain the tile issued n_streams ago: its counts have landed by now
            if len(pending) > len(self.streams):
                d2h += self._drain(pending.pop(0))
        while pending:
            d2h += self._drain(pending.pop(0))
        for s in self.streams:
            cur.wait_stream(s)
        cur.synchronize()
        for item in self._expanding:                       # tiles whose sparse copy was still in flight
            self._expand(*item)
        if self.pool is not None:
            for f in self._futures:
                f.result()
        res = []
        for o in self.out:
            c = o["counts"].tolist()
            v = {"dense": o["dense"][:c[0]], "T": o["T"], "status": o["status"], "median_resolution": o["median_resolution"]}
            if self.want_sparse:
                v["sparse"] = o["sparse"][:(2 * c[1] if self.sparse_once else c[1])]
            res.append(v)
        return res, h2d, d2h

    def _expand(self, i, ev, rows):
        ev.synchronize()
        o = self.out[i]
        self._futures.append(self.pool.submit(ops.host_expand_sparse, o["sparse_once"][:rows], o["pair_rows"], o["sparse"], 2))

    def _drain(self, item):
        i, st, ev, r, t = item
        ev.synchronize()
        o = self.out[i]
        c = o["counts"].tolist()
        n = 4 * 4 + 4 + o["T"].numel() * 4 + o["status"].numel()
        # the previous tile's sparse rows have landed by now: hand them to the host threads
        while self._expanding and self._expanding[0][1].query():
            self._expand(*self._expanding.pop(0))
        with torch.cuda.stream(st):
            o["dense"][:c[0]].copy_(r.dense[:c[0]], non_blocking=True)
            n += c[0] * 24
            if self.sparse_once:
                o["sparse_once"][:c[1]].copy_(r.sparse[:c[1]], non_blocking=True)
                n += c[1] * 24 + o["pair_rows"].numel() * 4
                ev2 = torch.cuda.Event()
                ev2.record(st)
                self._expanding.append((i, ev2, c[1]))
            elif self.want_sparse:
                o["sparse"][:c[1]].copy_(r.sparse[:c[1]], non_blocking=True)
                n += c[1] * 24
            # keep the device tensors alive until the copies are done
            r.dense.record_stream(st)
            r.sparse.record_stream(st)
        return n


def f2s3_tile(src, tgt, feat_src, feat_tgt, svl_ptr, svl_idx, weights=None, filter_net=None, coeff=1.0,
              refine_results=False, max_disp_magnitude=5.0, mutual=False, want_median=True):
    """The F2S3 hot path on one tile, device -> device (BASELINE config C2; src/f2s3.py:248-441 without the files):
    A1 median resolution (the descriptor radius of compute_features, f2s3.py:106) -> B1 exact descriptor 1-NN on the
    tensor cores -> rows [src | tgt[label]] in supervoxel order (CSR svl_ptr / svl_idx over source points) -> weights
    (filter_net.compute_weights_segments, or `weights` (n,) handed in) -> per supervoxel weighted Kabsch -> residual
    median filter -> refit (D3 F4) -> keep mask -> magnitude gate (F1).
    mutual: both search directions; rows whose target's nearest source is another point get weight 0.
    Returns a dict of device tensors; `rows` / `mag` are the kept correspondences (one host sync for their count)."""
    from . import f2s3 as hot
    out = {}
    if want_median:
        out["median_resolution"] = ops.median_resolution(src, tgt)
    if mutual:
        ri, _, ci, _ = ops.desc_nn(feat_src, feat_tgt, both_dirs=True)
    else:
        ri, _ = ops.desc_nn(feat_src, feat_tgt)
        ci = None
    out["labels"] = ri
    sel = svl_idx.long()
    lab_sel = ri[sel].long()
    X = torch.cat([src[sel], tgt[lab_sel]], dim=1).contiguous()                      # f2s3.py:284-285, :366
    ptr = svl_ptr.to(src.device, torch.int32).contiguous()
    if filter_net is not None:
        with torch.no_grad():
            scores = filter_net.compute_weights_segments(X, ptr, scale=True)
    else:
        scores = weights.to(torch.float32)[sel]
    if mutual:
        scores = torch.where(ci[lab_sel] == svl_idx.to(ci.dtype), scores, torch.zeros_like(scores))
    scores = scores.contiguous()
    R, t, robust, res = hot.filter_input_tail(X, scores, ptr, coeff)
    keep = scores > 0.99999
    if refine_results:
        seg = torch.repeat_interleave(torch.arange(ptr.numel() - 1, device=src.device), (ptr[1:] - ptr[:-1]).long())
        keep = keep | robust[seg]
    rows = X[keep].contiguous()
    if rows.shape[0] > 0:                                        # f2s3.py:392-393: `<= max`, unconditional
        mask, mag = ops.magnitude_mask(rows, max_mag=float(max_disp_magnitude))
        m = mask.bool()
        rows, mag = rows[m], mag[m]
    else:
        mag = torch.zeros((0,), dtype=torch.float32, device=rows.device)
    out.update(rows=rows, mag=mag, keep=keep, scores=scores, R=R, t=t, robust=robust, residuals=res, corr=X)
    return out
