#!/usr/bin/env python
"""Headline benchmark: displacement-field points/sec of the patch-wise hot path.

Workload (BASELINE.json configs[4], "C5"): a 50 M-point epoch pair as 64 synthetic tiles of
781 250 points per epoch, ~256-point patches, point correspondences resident as (N,2) int64 --
the shape fusion4landslide processes tile by tile.  One step = one pass of the hot path over all
tiles: A1 median resolution (k=2 self-kNN of both epochs) and the fused fine-matching stage
(F2 select, F3 rigidity, D2 Procrustes, E1 per-patch ICP, D5 apply -> dense DVF, A4 1-NN assign ->
sparse DVF).  Tiles are independent: with N GPUs they are dealt round-robin (equal tiles, LPT is a
no-op) and the per-patch transforms and the dense DVF are all-gathered over NCCL at the end of the
step ("strong" scaling: the 50 M-point job is fixed).

    python bench.py --gpus N --steps K --warmup W          (torchrun for N > 1)
    python bench.py --impl reference ...                   the CPU arm (oracle port, all host cores)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# more hardware work queues than the default 8: the step uses 4 tile streams + 4 side streams (inside a CUDA graph), NCCL's
# stream and the exchange stream; with 8 queues the exchange stream shares one with graph work and serialises behind it
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "displacement_field_points_per_sec"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiles", type=int, default=64)
    ap.add_argument("--tile-pts", type=int, default=781_250)
    ap.add_argument("--patch-pts", type=int, default=256)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--compact-corr", type=int, default=0,
                    help="e2e: host threads repack the int64 (n,2) correspondence tables to the int32 column the stage reads "
                         "(inside the timed region); 4 instead of 16 bytes per source point are uploaded")
    ap.add_argument("--c2f-threads", type=int, default=2, help="C3 / C4: host threads (one CUDA stream each) the tiles are spread over")
    ap.add_argument("--streams", type=int, default=4, help="side streams for tile-level concurrency (1 = serial)")
    ap.add_argument("--workload", default="c5", choices=["c5", "dips"],
                    help="c5: the headline hot path (default).  dips: the DIPs patch front-end of SURVEY 8(f) rank 1 "
                         "(one 625 k-point tile-epoch per step) with its own metric and CPU leg")
    ap.add_argument("--dips-pts", type=int, default=625_000)
    ap.add_argument("--dips-cpu-queries", type=int, default=300)
    ap.add_argument("--a1-overlap", type=int, default=1, help="1: A1 of a tile runs on a side stream next to its rigid fits")
    ap.add_argument("--fit", default="per-tile", choices=["batched", "per-tile"],
                    help="per-tile (default): one fit launch per tile, tiles overlapped on streams.  batched: the rigid fits of "
                         "a group of tiles in ONE persistent queue-driven launch (pipeline.displacement_field_tiles_batched): "
                         "the fit kernel alone is 21 %% faster (no wave quantisation), the step 13 %% slower (it fills every "
                         "register and leaves the other phases no tail to overlap with) -- measured, DESIGN.md section 4")
    ap.add_argument("--fit-groups", type=int, default=1, help="batched fit: tile groups per rank (one fit launch each)")
    ap.add_argument("--fit-ctas", type=int, default=0, help="batched fit: CTAs per SM of the persistent kernel (0 = 4)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--no-gather", action="store_true", help="diagnosis only: skip the exchange (N > 1)")
    ap.add_argument("--exchange", default="push", choices=["push", "fused", "nccl"],
                    help="N > 1: 'push' (default) = a copy kernel per tile (TMA bulk: local rows -> shared memory -> every "
                         "peer's field over NVLink) on an exchange stream next to the following tiles' fits; 'fused' = the D5 "
                         "kernel itself stores every dense row into all peers' fields; 'nccl' = all_gather_into_tensor after "
                         "the step (baseline)")
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="patch pairs in the CPU sample (0 = auto)")
    ap.add_argument("--scene", default="v2", choices=["v1", "v2"],
                    help="v2 (default): SURVEY 8(d) scenes -- epoch 2 an INDEPENDENT resample, nested patch hierarchy, "
                         "~2 %% of the points in patches <= 10 points.  v1: round-1 scenes (epoch 2 = jittered epoch 1)")
    ap.add_argument("--configs", default="C1,C2,C3,C4",
                    help="BASELINE.json configs measured next to the headline workload on rank 0 (N = 1 only): "
                         "comma list out of C1,C2,C3,C4, or 'none'")
    ap.add_argument("--config-steps", type=int, default=2)
    return ap.parse_args()


def workload_name(a):
    return ("C5: %d tiles x %d pts/epoch (%.1fM-point epoch pair), ~%d-pt patches, fusion4landslide fine-matching "
            "path (A1 median-resolution kNN + F2/F3/D2/E1/D5/A4) per tile; scene %s" %
            (a.tiles, a.tile_pts, a.tiles * a.tile_pts / 1e6, a.patch_pts,
             "v2 (independent epochs, SURVEY 8d)" if a.scene == "v2" else "v1 (epoch 2 = jittered epoch 1)"))


def make_c5_tile(a, seed, dev):
    from fusion4landslide_b200 import synth
    if a.scene == "v1":
        return synth.make_tile(a.tile_pts, seed=seed, device=dev, patch_pts=a.patch_pts, origin=(0.0, 0.0))
    return synth.make_scene(a.tile_pts, seed=seed, device=dev)          # label_src / label_tgt: ~256-point patches


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                power.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full summary under profiles/
    (dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_%s.txt" % kernel))):
        tot, unit_mult = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        found = 0
        for line in open(path):
            m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if m and found < 2:
                tot += float(m.group(2)) * unit_mult.get(m.group(3), 1.0)
                found += 1
        if found == 2:
            best = (tot, os.path.basename(path))
    return best


def ncu_metrics(kernel, names):
    """Selected metrics of `kernel` from the newest committed ncu --set full summary under profiles/ (context for a
    kernel that is not memory-bound: issue-slot and pipe utilisation), {} when there is none."""
    import glob
    import re
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_%s.txt" % kernel))):
        found = {}
        for line in open(path):
            m = re.match(r"\s*(\S+)\s+([0-9.]+)\s*(\S*)", line)
            if m and m.group(1) in names and m.group(1) not in found:
                found[m.group(1)] = float(m.group(2))
        if found:
            out = dict(found, source=os.path.basename(path))
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# algorithmic bytes per launch of each kernel, given the tile quantities (DESIGN.md "Kernels")
def algorithmic_bytes(name, q):
    N, M, P = q["n_src"], q["n_tgt"], q["pairs"]
    ns, nt, K = q["src_items"], q["tgt_items"], q["matched"]
    table = {
        # A1 (medres.cu) handles BOTH epochs per launch.  search: every float4 row once as reference and once as
        # query (12N + 12M + 8kN with N = M per epoch, k = 2 -> 40 B/point; the kernel moves 16 + 4)
        "k_a1_search": 40 * (N + M),
        "k_a1_search_tiled": 40 * (N + M),
        "k_a1_bbox": 12 * (N + M),
        "k_a1_count": 12 * (N + M),
        "k_a1_scatter": (12 + 16) * (N + M),
        "k_a1_select": 4 * (N + M),
        # corr col1 (8 B) + labels (4 B) + point index (4 B) per src patch item, 8 B per selected pair
        "k_select_corr": 16 * ns + 8 * K,
        # matched pairs: 8 B indices + 24 B coordinates, outputs 64+128+... per pair
        "k_patch_fit": 32 * K + 240 * P,
        "k_patch_fit_warp": 32 * K + 240 * P,
        "k_patch_fit_warp_tiles": 32 * q.get("matched_all", K) + 240 * q.get("pairs_all", P),     # one launch, all tiles of the rank
        # src patch points 12+4, tgt patch points 12+4, dense rows 24, nn 4
        "k_apply_assign": 16 * ns + 16 * nt + 24 * ns + 4 * ns + 64 * P,
        "k_emit_sparse": 8 * ns + 2 * 24 * q["sparse_half"] + 24 * q["sparse_half"],
    }
    return table.get(name)


# ---------------------------------------------------------------------------------------------
def run_b200(a):
    import torch.distributed as dist
    from fusion4landslide_b200 import _lib, pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"            # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    cfg = pipeline.FineConfig()

    my_tiles = list(range(rank, a.tiles, world))           # equal tiles: round-robin == LPT
    tiles, host_tiles = [], []
    for t in my_tiles:
        d = make_c5_tile(a, a.seed * 100003 + t, dev)
        tiles.append(pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"]))
        del d
    torch.cuda.synchronize()

    # per-rank output arena: every tile writes its dense rows / transforms into its slice, the
    # arena is what the all-gather ships
    cap_rows = sum(t.n_src_items for t in tiles)
    cap_pairs = sum(t.n_pairs for t in tiles)
    if world > 1:
        caps = torch.tensor([cap_rows, cap_pairs], device=dev, dtype=torch.int64)
        dist.all_reduce(caps, op=dist.ReduceOp.MAX)
        cap_rows, cap_pairs = int(caps[0]), int(caps[1])
    fused = world > 1 and a.exchange in ("fused", "push") and not a.no_gather      # dense rows travel through peer memory
    pushing = fused and a.exchange == "push"
    ex = None
    if fused:
        from fusion4landslide_b200.exchange import PeerExchange
        ex = PeerExchange(cap_rows, dev, n_buffers=2)        # double-buffered by step parity
        arenas = [ex.local_arena(0), ex.local_arena(1)]
    else:
        arenas = [torch.empty((cap_rows, 6), dtype=torch.float32, device=dev)]
    dense_arena = arenas[0]
    # per parity ONE small arena [transforms | row counts]: a single all-gather per step ships both, and the step that
    # overwrites it is two steps later (the gather of step k runs on the exchange stream under step k + 1)
    n_tiles_max = (a.tiles + world - 1) // world
    meta_arenas = [torch.zeros((cap_pairs * 16 + n_tiles_max * 4,), dtype=torch.float32, device=dev) for _ in arenas]
    T_arenas = [m[:cap_pairs * 16].view(cap_pairs, 4, 4) for m in meta_arenas]
    counts_arenas = [m[cap_pairs * 16:].view(torch.int32).view(n_tiles_max, 4) for m in meta_arenas]
    T_arena, counts_arena = T_arenas[0], counts_arenas[0]
    meds = torch.empty((max(len(tiles), 1),), dtype=torch.float32, device=dev)
    from fusion4landslide_b200 import ops as ops_mod
    from fusion4landslide_b200.ops import FineResult
    outs_par, peers_par = [], []
    shared = None
    for par, arena in enumerate(arenas):
        outs, peers = [], []
        ro = po = 0
        for i, t in enumerate(tiles):
            r = FineResult()
            Q = t.n_pairs
            if par == 0:
                r.T = T_arena[po:po + Q]
                r.T64 = torch.empty((Q, 4, 4), dtype=torch.float64, device=dev)
                r.status = torch.empty((Q,), dtype=torch.int8, device=dev)
                r.K = torch.empty((Q,), dtype=torch.int32, device=dev)
                r.fitness = torch.empty((Q,), dtype=torch.float64, device=dev)
                r.rmse = torch.empty((Q,), dtype=torch.float64, device=dev)
                r.iters = torch.empty((Q,), dtype=torch.int32, device=dev)
                r.ratio_inlier = torch.empty((Q,), dtype=torch.float32, device=dev)
                r.dist_mean = torch.empty((Q,), dtype=torch.float32, device=dev)
                r.sparse = torch.empty((2 * t.n_src_items, 6), dtype=torch.float32, device=dev)
                r.tgt2src = None
                r.sparse_pair_rows = None
                r.icp_fragile = torch.zeros((Q,), dtype=torch.uint8, device=dev)
                r.counts = counts_arena[i]
            else:                                           # parity 1 differs only in the dense arena
                for k in FineResult.__slots__:
                    setattr(r, k, getattr(outs_par[0][i], k, None))
            if par > 0:                                     # the copy kernels / the gather of step k run while step k+1 computes
                r.counts = counts_arenas[par][i]
                r.T = T_arenas[par][po:po + Q]
            r.dense = arena[ro:ro + t.n_src_items]
            outs.append(r)
            peers.append(ex.peer_ptrs(par, ro) if fused else None)
            ro += t.n_src_items
            po += Q
        outs_par.append(outs)
        peers_par.append(peers if fused else None)
    outs = outs_par[0]
    if world > 1:
        gathered_dense = None if fused else torch.empty((world * cap_rows, 6), dtype=torch.float32, device=dev)
        gathered_meta = torch.empty((world, meta_arenas[0].numel()), dtype=torch.float32, device=dev)
        gathered_counts = gathered_meta[:, cap_pairs * 16:]            # (world, n_tiles_max * 4) int32 bit patterns

    streams = pipeline.make_streams(a.streams, dev) if a.streams > 1 else None
    sides = pipeline.make_streams(a.streams, dev) if (a.streams > 1 and a.a1_overlap) else None

    caches = [{} for _ in arenas]
    fit_stream = torch.cuda.Stream(device=dev) if (streams and a.fit == "batched") else None
    # high priority: the copy kernels are a handful of one-warp CTAs that must not queue behind the fit kernels' waves
    xstream = torch.cuda.Stream(device=dev, priority=-1) if pushing else None

    def step_tiles(par=0):
        if a.fit == "batched":
            pipeline.displacement_field_tiles_batched(tiles, cfg, outs_par[par], meds, streams, peers_par[par],
                                                      side_streams=sides, cache=caches[par], groups=a.fit_groups,
                                                      fit_stream=fit_stream, fit_ctas_per_sm=a.fit_ctas)
        elif pushing:                                   # the copy kernels are enqueued by step(), after the graph
            pipeline.displacement_field_tiles(tiles, cfg, outs_par[par], meds, streams, None, side_streams=sides)
        else:
            pipeline.displacement_field_tiles(tiles, cfg, outs_par[par], meds, streams, peers_par[par], side_streams=sides)

    # The per-tile launch sequence is static (all buffers preallocated, no host round trip inside the path):
    # capture one step -- all tiles, all side streams -- into a CUDA graph and replay it (one graph per
    # exchange-buffer parity).
    graphs = None
    if not a.no_graph:
        step_tiles(0)                                  # sizes the per-stream workspaces outside the capture
        torch.cuda.synchronize()
        try:
            graphs = []
            for par in range(len(arenas)):
                cap = torch.cuda.Stream(device=dev)
                cap.wait_stream(torch.cuda.current_stream(dev))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap):
                    step_tiles(par)
                torch.cuda.current_stream(dev).wait_stream(cap)
                graphs.append(g)
        except Exception as e:                         # report, then fall back to plain launches
            sys.stderr.write("bench.py: CUDA graph capture failed (%s); launching kernel by kernel\n" % (e,))
            graphs = None
            torch.cuda.synchronize()
    graph = graphs
    step_no = [0]

    def compute(par):
        if graphs is not None:
            graphs[par].replay()
        else:
            step_tiles(par)

    def exchange(par=0):
        # fused: the dense rows are already in every rank's field; this small all-gather is also the barrier
        # that orders all ranks' pushed rows before anyone reads the field
        if not fused:
            dist.all_gather_into_tensor(gathered_dense, dense_arena)
        dist.all_gather_into_tensor(gathered_meta.view(-1), meta_arenas[par % len(meta_arenas)])

    push_done = [None]

    def launch_pushes(par):
        """Exchange of the step just computed, pipelined with the NEXT step: one copy kernel per tile on the exchange
        stream (ops.peer_push: local rows -> shared memory -> every peer's field, TMA bulk copies).  Half of a rank's
        rows are produced by its last wave of tiles, so an exchange that must end with the step exposes ~0.7 ms of
        NVLink time at 8 GPUs (measured: 5.2 ms per step fused or per-tile pushed, 4.1 ms without any exchange); the
        fields are double-buffered by step parity, so step k's copies run under step k+1's fits instead."""
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(xstream):
            xstream.wait_event(ev)
            for i in range(len(tiles)):
                ops_mod.peer_push(outs_par[par][i].dense, outs_par[par][i].counts, peers_par[par][i])
            # the small all-gather (transforms + counts) follows the copies ON THE EXCHANGE STREAM: when it completes
            # every rank has issued its copies of this step, and the main stream never waits for a collective
            exchange(par)
            done = torch.cuda.Event()
            done.record(xstream)
        push_done[0] = done

    def drain_pushes():
        if push_done[0] is not None:
            torch.cuda.current_stream(dev).wait_event(push_done[0])
            push_done[0] = None

    def step():
        par = step_no[0] % len(arenas)
        step_no[0] += 1
        compute(par)
        if world > 1 and not a.no_gather:
            if pushing:
                drain_pushes()             # the previous step's rows have left and its gather has passed on every rank
                launch_pushes(par)
            else:
                exchange(par)

    def barrier():
        torch.cuda.synchronize()           # own copy kernels first: a peer may read its field right after the barrier
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    drain_pushes()
    barrier()
    L.f4l_launch_count_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    drain_pushes()                         # the last step's exchange ends inside the timed region
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = L.f4l_launch_count()
    if graph is not None:
        # replayed kernels do not pass through the library's host-side counter: count one uncaptured step
        L.f4l_launch_count_reset()
        step_tiles()
        torch.cuda.synchronize()
        launches = L.f4l_launch_count() * a.steps
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms[0]) / a.steps

    # units processed: source points that received a displacement vector (dense DVF rows), all ranks
    rows = counts_arena[:, 0].sum().to(torch.int64).reshape(1)
    stat = torch.stack([rows[0], torch.tensor(sum(t.src.shape[0] for t in tiles), device=dev),
                        torch.tensor(launches, device=dev)]).to(torch.int64)
    if world > 1:
        dist.all_reduce(stat)
    dvf_points, src_points, launches_all = int(stat[0]), int(stat[1]), int(stat[2])
    value = dvf_points / (ms_step * 1e-3)

    # ---- breakdown (N > 1): one extra step with the compute and the exchange timed apart ---------
    breakdown = None
    if world > 1 and not a.no_gather:
        ea, eb, ec = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier()
        ea.record()
        par_b = step_no[0] % len(arenas)
        compute(par_b)
        step_no[0] += 1
        eb.record()
        if pushing:
            launch_pushes(par_b)
            drain_pushes()                 # un-pipelined here: the copies' own duration shows up in exchange_ms
        else:
            exchange(par_b)
        ec.record()
        barrier()
        bd = torch.tensor([ea.elapsed_time(eb), eb.elapsed_time(ec)], device=dev, dtype=torch.float64)
        dist.all_reduce(bd, op=dist.ReduceOp.MAX)
        breakdown = {"compute_ms": float(bd[0]), "exchange_ms": float(bd[1]), "exchange": a.exchange,
                     "exchange_bytes_received_per_rank": int((world - 1) * (cap_rows * 24 + cap_pairs * 64))}
        if fused:
            # check the fused exchange against NCCL: gather the local arenas of the last step the plain way
            last = (step_no[0] - 1) % len(arenas)
            ref = torch.empty((world * cap_rows, 6), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(ref, arenas[last])
            ref = ref.view(world, cap_rows, 6)
            cnt = gathered_counts.contiguous().view(torch.int32).view(world, n_tiles_max, 4)[:, :, 0].sum(1).tolist()
            ok = all(torch.equal(ex.field(last)[r, :cnt[r]], ref[r, :cnt[r]]) for r in range(world))
            okt = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            breakdown["fused_field_equals_nccl_all_gather"] = bool(int(okt[0]))
            del ref

    # ---- e2e: same step through the host-buffer API (H2D inputs + D2H results inside the timing) ----
    e2e = None
    if not a.no_e2e:
        host_tiles = [pipeline.HostTile(t) for t in tiles]
        hp = pipeline.HostPipeline(host_tiles, cfg, dev, n_streams=max(a.streams, 1), compact_corr=bool(a.compact_corr))
        hp.run()                                                            # warm-up (pinned buffers are allocated above)
        barrier()
        n_e2e = max(1, min(a.steps, 3))
        e0.record()
        h2d = d2h = 0
        rows_e2e = 0
        for _ in range(n_e2e):
            res, bi, bo = hp.run()
            h2d += bi
            d2h += bo
            rows_e2e += sum(r["dense"].shape[0] for r in res)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1) / n_e2e
        te = torch.tensor([ms_e2e, float(rows_e2e) / n_e2e, float(h2d) / n_e2e, float(d2h) / n_e2e], device=dev,
                          dtype=torch.float64)
        if world > 1:
            mx = te[:1].clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = te[1:].clone()
            dist.all_reduce(sm)
            te = torch.cat([mx, sm])
        e2e = {"value": float(te[1]) / (float(te[0]) * 1e-3), "unit": UNIT, "ms_per_step": float(te[0]),
               "steps": n_e2e, "h2d_bytes_per_step": int(te[2]), "d2h_bytes_per_step": int(te[3]),
               "note": ("pinned host inputs -> device, path, dense+sparse DVF/transforms -> pinned host; tiles pipelined over %d streams" % max(a.streams, 1)) +
                       ("; the int64 (n,2) correspondence tables are repacked to their int32 target column by host threads inside the timed region" if a.compact_corr else "")}
        del host_tiles, hp

    # ---- roofline leg: the same step with in-stream per-kernel CUDA events -----------------------
    roofline, kernel_table = None, None
    if rank == 0:
        L.f4l_profile_reset()
        L.f4l_profile_enable(1)
        prof_cache = {}
        for _ in range(max(1, min(a.steps, 3))):
            if a.fit == "batched":               # the timed step on ONE stream, so that in-stream events time each kernel alone
                pipeline.displacement_field_tiles_batched(tiles, cfg, outs, meds, None, None, None, cache=prof_cache)
            else:
                for i, t in enumerate(tiles):
                    pipeline.displacement_field(t, cfg, out=outs[i], med_out=meds[i:i + 1])
        torch.cuda.synchronize()
        L.f4l_profile_enable(0)
        del prof_cache
        tab = _lib.profile_table()
        tot = sum(v[0] for v in tab.values()) or 1.0
        kernel_table = {k: {"ms_avg": v[0] / v[1], "launches": v[1], "share": v[0] / tot}
                        for k, v in sorted(tab.items(), key=lambda kv: -kv[1][0])}
        top = next(iter(kernel_table))
        t0_ = tiles[0]
        c = counts_arena[0].tolist()
        q = dict(n_src=t0_.src.shape[0], n_tgt=t0_.tgt.shape[0], pairs=t0_.n_pairs, src_items=t0_.n_src_items,
                 tgt_items=t0_.n_tgt_items, matched=int(outs[0].K.sum()), sparse_half=c[1] // 2,
                 matched_all=int(sum(int(o.K.sum()) for o in outs)), pairs_all=sum(t.n_pairs for t in tiles))
        peak, peak_src = load_peaks()
        ab = algorithmic_bytes(top, q)
        ach = ab / (kernel_table[top]["ms_avg"] * 1e-3) / 1e9 if ab else None
        tr = ncu_traffic(top)
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": (ach / peak) if ach else None, "traffic": tr[0] if tr else None,
                    "traffic_source": tr[1] if tr else None,
                    "algorithmic_bytes_per_launch": ab, "ms_per_launch": kernel_table[top]["ms_avg"],
                    "share_of_step": kernel_table[top]["share"], "peak_source": peak_src,
                    "note": "kernel durations from in-stream CUDA events in a separate profiled (single-stream) pass of the "
                            "same step; this kernel is FP/issue-bound (K^2 rigidity + the on-chip ICP loop: ncu DRAM < 1 % of "
                            "peak, issue slots ~50 %), so its HBM fraction is small by construction -- see DESIGN.md section 4"}
        ctx = ncu_metrics(top, ("smsp__issue_active.avg.pct_of_peak_sustained_active",
                                "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                                "sm__warps_active.avg.pct_of_peak_sustained_active",
                                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                                "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                                "smsp__inst_executed.sum"))
        if ctx:
            roofline["ncu_context"] = ctx          # from the committed capture, not measured in this run
            inst = ctx.get("smsp__inst_executed.sum")
            if inst and top.startswith("k_patch_fit_warp"):
                # the roof this kernel actually sits under: warp-instruction issue slots (4 schedulers per SM, one
                # instruction per cycle each).  Instructions per launch from the committed ncu capture of the same kernel
                # on the same tile shape; duration measured in this run.
                sm_hz = 1.965e9
                peak = 148 * 4 * sm_hz
                ach = inst * (q["pairs_all"] / q["pairs"] if top.endswith("_tiles") else 1.0) / (kernel_table[top]["ms_avg"] * 1e-3)
                roofline["issue"] = {"bound": "issue", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp-inst/s",
                                     "frac": ach / peak, "warp_instructions_per_launch": inst,
                                     "note": "148 SMs x 4 schedulers x 1.965 GHz; instruction count from " + ctx.get("source", "ncu")}
        # also report the two kernels the north star names (kNN search, Kabsch/ICP reduction)
        for name in ("k_a1_search", "k_a1_scatter", "k_a1_count", "k_a1_bbox", "k_patch_fit_warp", "k_patch_fit_warp_tiles", "k_patch_fit",
                     "k_apply_assign"):
            if name in kernel_table:
                b = algorithmic_bytes(name, q)
                kernel_table[name]["hbm_frac"] = b / (kernel_table[name]["ms_avg"] * 1e-3) / 1e9 / peak

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ------------------
    input_gb = sum(t.nbytes() for t in tiles) * world / 1e9
    cpu, parity = None, None
    if rank == 0 and not a.no_cpu_baseline:
        raw = []
        cpu = cpu_arm_sample(a, tiles[0], cfg, target_seconds=15.0, raw_out=raw)
        parity = c5_parity(raw[0][0], raw[0][1], tiles[0], outs[0])

    # ---- BASELINE configs C1-C4 (rank 0, single GPU runs only) -----------------------------------
    configs_out = None
    if rank == 0 and world == 1 and a.configs.lower() != "none":
        used_graph = graph is not None
        del tiles, outs, outs_par, arenas, dense_arena, T_arena, graphs, graph, caches
        graph = True if used_graph else None
        torch.cuda.empty_cache()
        configs_out = run_configs(a, dev, L)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 i/o, f64 accumulation", "data": "synthetic",
            "config": {"workload": workload_name(a), "tiles": a.tiles, "tile_pts": a.tile_pts, "scene": a.scene,
                       "src_points_per_step": src_points, "dvf_points_per_step": dvf_points,
                       "parallelism": ("tile-sharded x%d; dense DVF rows copied into every GPU's field over NVLink by TMA copy kernels (one "
                                       "per tile, peer memory) on an exchange stream, pipelined with the next step's fits (fields "
                                       "double-buffered by step parity; the last step's copies end inside the timed region); "
                                       "transforms + row counts: one NCCL all-gather per step on the same exchange stream" % world) if pushing
                       else ("tile-sharded x%d; dense DVF rows stored into every GPU's field by the producing kernel over "
                             "NVLink (peer memory), transforms + row counts all-gathered over NCCL" % world) if fused
                       else "tile-sharded x%d, NCCL all-gather of transforms + dense DVF" % world,
                       "l2": "inputs (%.1f GB per step) exceed the 126 MB L2; no explicit flush" %
                             input_gb,
                       "icp_threshold": cfg.icp_threshold, "assign_type": cfg.assign_type, "streams": a.streams,
                       "cuda_graph": graph is not None, "a1_side_stream": bool(sides), "fit": a.fit, "fit_groups": a.fit_groups, "fit_ctas_per_sm": a.fit_ctas or 4},
            "e2e": e2e, "gpu_launches": launches_all, "clocks": clocks, "roofline": roofline,
            "kernels": kernel_table, "cpu_baseline": cpu, "parity": parity, "breakdown": breakdown,
            "configs": configs_out,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def c5_parity(r, spt_src, tile, out):
    """In-run parity of tile 0: the CPU arm's (oracle) per-pair results against the GPU's, reported in the bench line
    (tie / path-flip counts are part of the gate)."""
    import numpy as np
    k = r["pairs"]
    Kg, Sg, Ig = (x[:k].cpu().numpy() for x in (out.K, out.status, out.iters))
    ok = (r["status"] == 0) & (Sg == 0)
    same = ok & (Ig == r["iters"])
    Tg = out.T[:k].cpu().numpy().astype(np.float64)
    src = tile.src.cpu().numpy().astype(np.float64)
    worst = 0.0
    for q in np.nonzero(same)[0][:2000]:
        P = src[spt_src[q]]
        To = r["T"][q].astype(np.float64)
        worst = max(worst, float(np.abs((P @ Tg[q][:3, :3].T + Tg[q][:3, 3]) - (P @ To[:3, :3].T + To[:3, 3])).max()))
    rejected = int((r["status"] == 1).sum())
    frag = getattr(out, "icp_fragile", None)
    frag = frag[:k].cpu().numpy() if frag is not None else np.zeros(k, np.uint8)
    flip = ok & (Ig != r["iters"])
    return {"checked_against": "oracle/cpu_path.py on the CPU-sampled pairs of tile 0, in this run", "pairs_checked": int(k),
            "icp_fragile_pairs": int((ok & (frag != 0)).sum()), "icp_path_flips_unflagged": int((flip & (frag == 0)).sum()),
            "K_mismatch": int((Kg != r["K"]).sum()), "status_mismatch": int((Sg != r["status"]).sum()),
            "rigidity_rejected_pairs": rejected, "rigidity_flips": int(((Sg == 1) != (r["status"] == 1)).sum()),
            "fitted_pairs": int(ok.sum()), "icp_path_flips": int((ok & (Ig != r["iters"])).sum()),
            "icp_iterations_mean": float(r["iters"][ok].mean()) if ok.any() else 0.0,
            "max_abs_dvf_diff_m_same_path": worst, "tolerance_m": 1e-5}


def cpu_arm_sample(a, tile, cfg, target_seconds=15.0, workers=None, pairs=None, raw_out=None):
    """Time the oracle port on a bounded sample of tile 0 (host copy).  Returns the cpu_baseline dict."""
    from oracle import cpu_path
    src = tile.src.cpu().numpy()
    tgt = tile.tgt.cpu().numpy()
    corr = tile.corr3d.cpu().numpy()
    sp_ptr, sp_idx = tile.sp_ptr.cpu().numpy(), tile.sp_idx.cpu().numpy().astype("int64")
    tp_ptr, tp_idx = tile.tp_ptr.cpu().numpy(), tile.tp_idx.cpu().numpy().astype("int64")
    Q = tile.n_pairs
    spt_src = [sp_idx[sp_ptr[q]:sp_ptr[q + 1]] for q in range(Q)]
    spt_tgt = [tp_idx[tp_ptr[q]:tp_ptr[q + 1]] for q in range(Q)]
    workers = workers or os.cpu_count() or 1
    if pairs is None:
        pairs = a.cpu_sample_pairs or min(Q, max(64, int(target_seconds * workers * 150)))   # ~6 ms per pair per core
    kw = dict(mode=cfg.mode, remove_low_quality_patch_matches=cfg.remove_low_quality_patch_matches,
              num_min_matches_for_quality_check=cfg.num_min_matches_for_quality_check,
              thres_dist_diff=cfg.thres_dist_diff, thres_inlier_ratio=cfg.thres_inlier_ratio,
              num_min_fine_match=cfg.num_min_fine_match, icp_refine=cfg.icp_refine, assign_type=cfg.assign_type,
              output_tgt2src=cfg.output_tgt2src, icp_threshold=cfg.icp_threshold, icp_max_iter=cfg.icp_max_iter)
    r = cpu_path.run_tile(src, tgt, corr, spt_src, spt_tgt, workers=workers, max_pairs=pairs, **kw)
    if raw_out is not None:
        raw_out.append((r, spt_src))
    # the median-resolution kNN covers the whole tile, the fine matching only the sampled pairs:
    # scale the kNN time by the sampled fraction so both legs describe the same points
    frac = r["src_points"] / max(1, src.shape[0])
    secs = r["seconds_fine"] + r["seconds_median"] * frac
    return {"value": r["dense_rows"] / secs, "unit": UNIT, "cores": workers, "kind": "port",
            "sample": "%d of %d patch pairs of tile 0 (%d src points; %.1f s fine matching + %.1f s x %.3f median-"
                      "resolution kNN); oracle/cpu_path.py: numpy + scipy cKDTree restatement of base.py:3254-3438, "
                      "patch pairs spread over %d forked workers" %
                      (r["pairs"], Q, r["src_points"], r["seconds_fine"], r["seconds_median"], frac, workers),
            "seconds": secs}



# =================================================================================================
# BASELINE.json configs C1-C4 next to the headline (C5) workload: each returns a dict with its own value,
# e2e, roofline (dominant kernel by in-stream CUDA-event time) and CPU leg.  Rank 0, one GPU.
# =================================================================================================
def _peaks_full():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": float(d["hbm_gbs"]), "tf_burst": float(d["bf16_tflops"]),
                "tf_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "MEASURED_PEAKS.json"}
    return {"hbm": 6650.0, "tf_burst": 1650.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def _timed(fn, warmup, steps, dev):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps, out


def _profiled(fn, L, dev):
    """One extra pass with the library's in-stream per-kernel events: {kernel: {ms_total, launches, share}}."""
    from fusion4landslide_b200 import _lib
    L.f4l_profile_reset()
    L.f4l_profile_enable(1)
    L.f4l_launch_count_reset()
    fn()
    torch.cuda.synchronize(dev)
    L.f4l_profile_enable(0)
    launches = int(L.f4l_launch_count())
    tab = _lib.profile_table()
    tot = sum(v[0] for v in tab.values()) or 1.0
    kern = {k: {"ms_total": round(v[0], 4), "launches": v[1], "share": round(v[0] / tot, 4)}
            for k, v in sorted(tab.items(), key=lambda kv: -kv[1][0])}
    return kern, launches


def _roofline(kern, work, peaks, note=""):
    """work: {kernel name: ("hbm", bytes per step) | ("tensor", flop per step)} for the kernels that can dominate."""
    top = next(iter(kern))
    k = kern[top]
    r = {"kernel": top, "share_of_step": k["share"], "ms_per_step": k["ms_total"], "launches_per_step": k["launches"],
         "peak_source": peaks["source"], "note": note}
    w = work.get(top)
    if w is None:
        r.update(bound=None, achieved=None, peak=None, unit=None, frac=None, traffic=None)
        return r
    kind, amount = w
    sec = k["ms_total"] * 1e-3
    if kind == "tensor":
        ach = amount / sec / 1e12
        r.update(bound="tensor", achieved=ach, peak=peaks["tf_sustained"], unit="TFLOP/s", frac=ach / peaks["tf_sustained"],
                 frac_of_burst=ach / peaks["tf_burst"], algorithmic_flop_per_step=amount)
    else:
        ach = amount / sec / 1e9
        r.update(bound="hbm", achieved=ach, peak=peaks["hbm"], unit="GB/s", frac=ach / peaks["hbm"],
                 algorithmic_bytes_per_step=amount)
    tr = ncu_traffic(top)
    r["traffic"] = tr[0] if tr else None
    r["traffic_source"] = tr[1] if tr else None
    ctx = ncu_metrics(top, ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
                            "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                            "smsp__issue_active.avg.pct_of_peak_sustained_active"))
    if ctx:
        r["ncu_context"] = ctx
    return r


def _cpu_backends():
    """Which of the reference's third-party back-ends exist on this host: the CPU legs use
    them when importable and the exact restatements otherwise."""
    import importlib.util
    return {m: importlib.util.find_spec(m) is not None for m in ("open3d", "hnswlib", "faiss", "sklearn", "scipy")}


def _cpu_desc_nn_rate(fs, ft, rows, backends):
    """Seconds per source row of the descriptor search on the host, the way the reference does it:
    hnswlib HNSW (M=12, efC=300, efS=300, 16 threads; f2s3_brienz.yaml:43-46) when the wheel exists, else the
    reference's exact branch: torch.cdist + min in batches of 1024 (base.py:2783-2815, 'cdist_cpu'), all cores."""
    fs = fs[:rows].contiguous()
    if backends["hnswlib"]:
        import hnswlib
        t0 = time.perf_counter()
        idx = hnswlib.Index(space="l2", dim=ft.shape[1])
        idx.init_index(max_elements=ft.shape[0], ef_construction=300, M=12)
        idx.set_num_threads(16)
        idx.add_items(ft.numpy())
        t_build = time.perf_counter() - t0
        idx.set_ef(300)
        t0 = time.perf_counter()
        idx.knn_query(fs.numpy(), k=1)
        return (time.perf_counter() - t0) / rows, "hnswlib HNSW M=12 efC=300 efS=300 (+%.1f s index build per tile)" % t_build, t_build
    t0 = time.perf_counter()
    for i in range(0, rows, 1024):
        d = torch.cdist(fs[i:i + 1024], ft)
        d.min(dim=1)
    return (time.perf_counter() - t0) / rows, "torch.cdist + min, batches of 1024 (base.py:2783-2815), %d torch threads" % torch.get_num_threads(), 0.0


def bench_c1(a, dev, L, peaks):
    """C1: Piecewise ICP (main_piecewise_icp.py) on one synthetic 1 M-point tile pair with known block shift."""
    import numpy as np
    from fusion4landslide_b200 import ops, synth
    n = 1_000_000
    d = synth.make_scene(n, seed=a.seed + 11, device=dev)
    src64, tgt64 = d["src"].double().contiguous(), d["tgt"].double().contiguous()
    smax, npmin = 5.0, 10
    step = lambda: ops.piecewise_icp(src64, tgt64, smax, npmin)
    ms, out = _timed(step, max(a.warmup, 2), a.config_steps + 1, dev)
    counts = out[2].tolist()
    rows = counts[0]
    kern, launches = _profiled(step, L, dev)
    work = {name: ("hbm", b) for name, b in (("cub_radix_sort", 2 * 12 * 2 * n * 3), ("k_pw_codes", 2 * n * (24 + 12)),
                                              ("k_pw_emit", 2 * n * 24 + rows * 48 + rows * 8), ("k_pw_centroids", 2 * n * 28),
                                              ("k_pw_leaves", 2 * n * 12), ("k_pw_order", 2 * n * 8))}
    res = {"workload": "C1: piecewise ICP, 1 tile, N = M = %d points, smax %.0f m, number_points_min %d" % (n, smax, npmin),
           "metric": METRIC, "unit": UNIT, "value": rows / (ms * 1e-3), "ms_per_step": ms, "steps": a.config_steps + 1,
           "dvf_points_per_step": rows, "src_points_per_step": n, "dtype": "f64", "gpu_launches_per_step": launches,
           "cells": {"src": counts[2], "tgt": counts[3], "octree_depth": counts[4], "unstable": counts[5]},
           "kernels": dict(list(kern.items())[:6]),
           "roofline": _roofline(kern, work, peaks, "bytes: keys+values of the radix-sort passes / the points and rows a kernel streams")}
    # e2e: pinned host float64 clouds -> device -> path -> rows back on the host
    hs, ht = src64.cpu().pin_memory(), tgt64.cpu().pin_memory()
    host_rows = torch.empty((n + 8, 6), dtype=torch.float64).pin_memory()

    def e2e_step():
        s_, t_ = hs.to(dev, non_blocking=True), ht.to(dev, non_blocking=True)
        o = ops.piecewise_icp(s_, t_, smax, npmin)
        c = o[2].tolist()
        host_rows[:c[0]].copy_(o[0][:c[0]], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return c[0]
    ms_e, r_e = _timed(e2e_step, 1, a.config_steps, dev)
    res["e2e"] = {"value": r_e / (ms_e * 1e-3), "unit": UNIT, "ms_per_step": ms_e, "h2d_bytes_per_step": 2 * n * 24,
                  "d2h_bytes_per_step": r_e * 48 + 24}
    if not a.no_cpu_baseline:
        from oracle import piecewise as opw
        s_np, t_np = src64.cpu().numpy(), tgt64.cpu().numpy()
        t0 = time.perf_counter()
        o = opw.piecewise_icp(s_np, t_np, smax, npmin)
        sec = time.perf_counter() - t0
        g = out[0][:rows].cpu().numpy()
        same_rows = o["dvfs"].shape[0] == rows
        err = float(np.abs(g - o["dvfs"]).max()) if same_rows else None
        res["cpu_baseline"] = {"value": o["dvfs"].shape[0] / sec, "unit": UNIT, "cores": 1, "kind": "port", "seconds": sec,
                               "sample": "the whole 1 M-point tile; oracle/piecewise.py (vectorised numpy + cKDTree restatement of "
                                         "piecewise_icp.py:89-216; the reference walks an Open3D octree in Python)"}
        res["parity"] = {"rows_equal": bool(same_rows), "max_abs_row_diff_m": err}
    del d, src64, tgt64
    return res


def bench_c2(a, dev, L, peaks):
    """C2: F2S3-style descriptor mutual-NN matching + weighted SVD, 4 tiles x 1 M points per epoch, D = 32 and 64."""
    import numpy as np
    from fusion4landslide_b200 import ops, pipeline, synth
    n, n_tiles = 1_000_000, 4
    res = {"workload": "C2: F2S3 path, %d tiles x %d pts/epoch: A1 + exact descriptor NN both directions (mutual) + per-"
                       "supervoxel weighted Kabsch -> median filter -> refit + magnitude gate; weights in [0,1] handed in" % (n_tiles, n),
           "metric": METRIC, "unit": UNIT, "dtype": "fp16 tensor-core candidates + f64 re-rank (descriptor NN), f32/f64 (Kabsch)", "by_D": {}}
    backends = _cpu_backends()
    for D in (32, 64):
        tiles = []
        for t in range(n_tiles):
            d = synth.make_scene(n, seed=a.seed * 7919 + 100 + t, device=dev, desc_dim=D)
            _, ptr, idx, _ = ops.labels_to_csr(d["label_src"].contiguous(), 10)
            g = torch.Generator(device=dev).manual_seed(t)
            w = torch.rand(n, generator=g, device=dev)
            w = torch.where(torch.rand(n, generator=g, device=dev) < 0.5, torch.ones_like(w), w)
            tiles.append(dict(src=d["src"], tgt=d["tgt"], fs=d["src_feat"], ft=d["tgt_feat"], ptr=ptr, idx=idx, w=w,
                              lab=d["label_src"]))
            del d

        def step():
            outs = []
            for t in tiles:
                outs.append(pipeline.f2s3_tile(t["src"], t["tgt"], t["fs"], t["ft"], t["ptr"], t["idx"], weights=t["w"],
                                               refine_results=True, max_disp_magnitude=5.0, mutual=True))
            return outs
        ms, outs = _timed(step, 1, a.config_steps, dev)
        rows = sum(int(o["rows"].shape[0]) for o in outs)
        kern, launches = _profiled(step, L, dev)
        flop = 2.0 * n * n * D * 2 * n_tiles                       # both directions
        work = {"k_desc_nn_tc": ("tensor", flop)}
        _, _, tie0 = ops.desc_nn(tiles[0]["fs"], tiles[0]["ft"], tie_eps=1e-6)
        r = {"value": rows / (ms * 1e-3), "src_points_per_sec": n * n_tiles / (ms * 1e-3), "ms_per_step": ms,
             "ties": {"desc_nn_rows_within_1e-6": int(tie0.sum()), "of_rows": int(tie0.numel())},
             "steps": a.config_steps, "dvf_points_per_step": rows, "src_points_per_step": n * n_tiles,
             "gpu_launches_per_step": launches, "kernels": dict(list(kern.items())[:6]),
             "mutual_fraction": float(sum((o["scores"] > 0).float().mean().item() for o in outs) / n_tiles),
             "roofline": _roofline(kern, work, peaks, "useful FLOP = 2*N*M*D per direction (the kernel issues D+16 K-columns); "
                                                      "peak = cuBLAS bf16 sustained (kernel timed inside a long step)")}
        # e2e: pinned host inputs -> device -> path -> kept rows + per-supervoxel transforms back
        host = [{k: v.cpu().pin_memory() for k, v in t.items() if k != "lab"} for t in tiles[:1]]

        def e2e_step():
            h = host[0]
            tt = {k: v.to(dev, non_blocking=True) for k, v in h.items()}
            o = pipeline.f2s3_tile(tt["src"], tt["tgt"], tt["fs"], tt["ft"], tt["ptr"], tt["idx"], weights=tt["w"],
                                   refine_results=True, max_disp_magnitude=5.0, mutual=True)
            back = [o["rows"].cpu(), o["mag"].cpu(), o["R"].cpu(), o["t"].cpu()]
            return o["rows"].shape[0], sum(x.numel() * x.element_size() for x in back)
        ms_e, (r_e, d2h) = _timed(e2e_step, 1, max(1, a.config_steps - 1), dev)
        r["e2e"] = {"value": r_e / (ms_e * 1e-3), "unit": UNIT, "ms_per_step": ms_e, "tiles_per_step": 1,
                    "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host[0].values()), "d2h_bytes_per_step": d2h}
        if not a.no_cpu_baseline:
            from oracle import paths as opaths
            t0_ = tiles[0]
            fs_h, ft_h = t0_["fs"].cpu(), t0_["ft"].cpu()
            sample_rows = 1024
            per_row, how, t_build = _cpu_desc_nn_rate(fs_h, ft_h, sample_rows, backends)
            per_row_b, t_build_b = per_row, t_build            # the reverse search has the same shape (N = M)
            o = opaths.f2s3_tile(t0_["src"].cpu().numpy(), t0_["tgt"].cpu().numpy(), fs_h.numpy(), ft_h.numpy(),
                                 t0_["lab"].cpu().numpy(), t0_["w"].cpu().numpy(), refine_results=True, max_disp_magnitude=5.0,
                                 mutual=False, max_segments=300, labels_given=outs[0]["labels"].long().cpu().numpy())
            from oracle import knn as oknn
            t1 = time.perf_counter()
            oknn.median_resolution(t0_["src"].cpu().numpy()[:250_000], t0_["tgt"].cpu().numpy()[:250_000])
            t_med = (time.perf_counter() - t1) * 4.0
            prune_rate = o["seconds"]["pruning"] / max(1, o["seconds"]["pruning_rows"])
            n_seg_rows = int(t0_["idx"].numel())
            sec_tile = (per_row + per_row_b) * n + t_build + t_build_b + prune_rate * n_seg_rows + t_med
            r["cpu_baseline"] = {"value": (rows / n_tiles) / sec_tile, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                 "seconds_per_tile_estimate": sec_tile,
                                 "sample": "tile 0, stage by stage, scaled to the tile: descriptor search %d rows x 1 M, counted for both directions "
                                           "(%s): %.2e s/row; pruning of 300 supervoxels (oracle/rigid.py filter_input_tail): %.2e s/row; "
                                           "median resolution on 250 k points x 4" % (sample_rows, how, per_row, prune_rate),
                                 "backends_present": backends}
        res["by_D"]["D%d" % D] = r
        del tiles, outs
        torch.cuda.empty_cache()
    # headline of the config = D = 32 (BASELINE.json configs[1]); D = 64 beside it
    for k in ("value", "ms_per_step", "steps", "dvf_points_per_step", "src_points_per_step", "e2e", "roofline", "cpu_baseline",
              "gpu_launches_per_step"):
        if k in res["by_D"]["D32"]:
            res[k] = res["by_D"]["D32"][k]
    return res


def _c2f_models(dev):
    """Random-init superpoint aggregation network of the reference's architecture (64-64-64; no checkpoint offline...
    the shipped parameters travel with tests/golden/nets_shipped.npz and are used when present)."""
    import numpy as np
    from fusion4landslide_b200 import nets
    m = nets.ClusterFeatureNetWithAttention()
    p = os.path.join(ROOT, "tests", "golden", "nets_shipped.npz")
    w = None
    if os.path.exists(p):
        z = np.load(p)
        w = {k[4:]: z[k] for k in z.files if k.startswith("agg/")}
        m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    else:
        torch.manual_seed(0)
        w = {k: v.numpy() for k, v in m.state_dict().items()}
    return m.to(dev).eval(), w


def bench_c2f(a, dev, L, peaks, fusion):
    """C3 (fusion_3d: coarse-to-fine on 3 superpoint levels, 16 tiles x 625 k) / C4 (C3 + 2D-lifted matches: 2D-vote
    coarse pairs and the lifted correspondences in the fine stage), through the CLASS-LEVEL entry point
    Coarse2Fine(cfg).implement_c2f_matching() on in-memory tiles."""
    import numpy as np
    from fusion4landslide_b200 import configs, synth
    from fusion4landslide_b200.entry_c2f import Coarse2Fine
    n, n_tiles, D = 625_000, 16, 64
    mode = "fusion" if fusion else "only_3d"
    model, w_np = _c2f_models(dev)
    tiles = []
    for t in range(n_tiles):
        d = synth.make_scene(n, seed=a.seed * 104729 + 300 + t, device=dev, desc_dim=D, frac_2d=0.05 if fusion else 0.0)
        tt = dict(src_pts=d["src"], tgt_pts=d["tgt"], partition_src=[d["labels_src"][k] for k in (1, 2, 3)],
                  partition_tgt=[d["labels_tgt"][k] for k in (1, 2, 3)], feat_raw_src=d["src_feat"], feat_raw_tgt=d["tgt_feat"])
        if fusion:
            tt["corres_3d_from_2d_idx"] = d["corr2d"]
        tiles.append(tt)
        del d
    last = {}

    def run_tile(tt):
        c = Coarse2Fine(configs.fusion_config(tt, mode=mode, levels=[1, 2, 3], feat_aggregate_model=model, device=str(dev)))
        c.implement_c2f_matching()
        return c

    # Tiles are independent: `--c2f-threads` host threads take them round-robin, each on its own CUDA stream, so the host side
    # of one tile (list / mask bookkeeping of the class path, a few row-count syncs) overlaps the kernels of another
    n_thr = max(1, min(int(a.c2f_threads), n_tiles))
    wstreams = [torch.cuda.Stream(device=dev) for _ in range(n_thr)] if n_thr > 1 else []
    pool = None
    if n_thr > 1:
        import concurrent.futures
        pool = concurrent.futures.ThreadPoolExecutor(max_workers=n_thr)

    def worker(w, cur):
        torch.cuda.set_device(dev)
        rows, c = 0, None
        with torch.cuda.stream(wstreams[w]):
            wstreams[w].wait_stream(cur)
            for tt in tiles[w::n_thr]:
                c = run_tile(tt)
                rows += int(c.data_output.corres_3d_refine_apply_icp.shape[0])
        return rows, c

    def step():
        rows = 0
        if n_thr > 1:
            cur = torch.cuda.current_stream(dev)
            res_w = [f.result() for f in [pool.submit(worker, w, cur) for w in range(n_thr)]]
            for s_ in wstreams:
                cur.wait_stream(s_)
            rows = sum(r_[0] for r_ in res_w)
            last["c"] = res_w[(n_tiles - 1) % n_thr][1]        # the worker that ran tiles[-1]
            return rows
        for tt in tiles:
            c = run_tile(tt)
            rows += int(c.data_output.corres_3d_refine_apply_icp.shape[0])
            last["c"] = c
        return rows
    ms, rows = _timed(step, 1, a.config_steps, dev)
    if pool is not None:
        pool.shutdown()
    kern, launches = _profiled(lambda: run_tile(tiles[0]), L, dev)
    c = last["c"]
    n_sub = int(c.data_interim.src_pts_sub.shape[0]), int(c.data_interim.tgt_pts_sub.shape[0])
    work = {"k_desc_nn_tc": ("tensor", 2.0 * n_sub[0] * n_sub[1] * D)}
    name = "C4" if fusion else "C3"
    res = {"workload": "%s: fusion_3d coarse-to-fine, %d tiles x %d pts/epoch, 3 superpoint levels (~64/192/576 pts)%s; per tile: voxel "
                       "subsampling (A1 A2) -> patch tables (F6) -> exact descriptor NN + gate + scatter (B2) -> per level attention "
                       "pooling (8f-2), mutual coarse NN (B3)%s, fused fine matching (F2 F3 D2 E1 D5 A4) -> level merge (M1); "
                       "through Coarse2Fine(cfg).implement_c2f_matching()" %
                       (name, n_tiles, n, " + 2D-lifted matches of 5 % of the source points" if fusion else "",
                        " + 2D vote (B4)" if fusion else ""),
           "metric": METRIC, "unit": UNIT, "value": rows / (ms * 1e-3), "src_points_per_sec": n * n_tiles / (ms * 1e-3),
           "ms_per_step": ms, "steps": a.config_steps, "dvf_points_per_step": rows, "src_points_per_step": n * n_tiles,
           "host_threads": n_thr,
           "dtype": "fp16 tensor-core candidates + f64 re-rank (descriptor NN), f32 i/o + f64 accumulation (fits)",
           "gpu_launches_per_tile": launches, "voxels_per_tile": list(n_sub),
           "pairs_per_level_last_tile": [len(x) for x in c.data_output.spt_corres_src_multiple],
           "kernels_tile0": dict(list(kern.items())[:8]),
           "roofline": _roofline(kern, work, peaks, "per tile (profiled pass over tile 0); useful FLOP = 2*N_sub*M_sub*D")}
    # e2e: the same call with HOST tensors (pinned) in tile_tensors, merged DVF rows back on the host
    host = {k: ([x.cpu().pin_memory() for x in v] if isinstance(v, list) else v.cpu().pin_memory()) for k, v in tiles[0].items()}

    def e2e_step():
        c = run_tile(host)
        do = c.data_output
        back = [do.corres_3d_refine_apply_icp.cpu(), do.corres_3d_refine_apply_icp_discrete.cpu()]
        return int(back[0].shape[0]), sum(x.numel() * x.element_size() for x in back)
    ms_e, (r_e, d2h) = _timed(e2e_step, 1, max(1, a.config_steps - 1), dev)
    h2d = sum(sum(x.numel() * x.element_size() for x in v) if isinstance(v, list) else v.numel() * v.element_size()
              for v in host.values())
    res["e2e"] = {"value": r_e / (ms_e * 1e-3), "unit": UNIT, "ms_per_step": ms_e, "tiles_per_step": 1,
                  "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    if not a.no_cpu_baseline:
        from oracle import paths as opaths
        backends = _cpu_backends()
        tt = tiles[-1]                                          # the tile `c` holds the GPU results of
        di = c.data_interim
        fs_sub, ft_sub = di.tile_pts_sub_feat_src.cpu(), di.tile_pts_sub_feat_tgt.cpu()
        sample_rows = 1024
        per_row, how, t_build = _cpu_desc_nn_rate(fs_sub, ft_sub, sample_rows, backends)
        pairs_cap = 60
        o = opaths.c2f_tile(tt["src_pts"].cpu().numpy(), tt["tgt_pts"].cpu().numpy(), [x.cpu().numpy() for x in tt["partition_src"]],
                            [x.cpu().numpy() for x in tt["partition_tgt"]], tt["feat_raw_src"].cpu().numpy(),
                            tt["feat_raw_tgt"].cpu().numpy(), w_np, voxel_size=float(c.method.voxel_size),
                            corr2d=tt["corres_3d_from_2d_idx"].cpu().numpy() if fusion else None, coarse=mode, fine=mode,
                            max_pairs_per_level=pairs_cap, corr3d_given=di.corres_3d_voxel_from_3d_idx.cpu().numpy(),
                            v2p_given={k: di["idx_voxel2pts_" + k].cpu().numpy() for k in ("src", "tgt")})
        sec = o["seconds"]
        pairs_all = sum(len(x) for x in c.data_output.spt_corres_src_multiple)
        pairs_done = sum(len(l["m"]) for l in o["levels"])
        pts_done = sum(l["n_src_points"] for l in o["levels"])
        pts_all = sum(int(sum(x.numel() for x in lv)) for lv in c.data_output.spt_corres_src_multiple)
        fine_full = sec["fine"] * pts_all / max(1, pts_done)
        sec_tile = (sec["voxel_subsampling"] + sec["median_resolution"] + per_row * n_sub[0] + t_build + sec["pooling"] +
                    sec["coarse"] + fine_full + sec["merge"])
        rows_tile = rows / n_tiles
        res["cpu_baseline"] = {"value": rows_tile / sec_tile, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                               "seconds_per_tile_estimate": sec_tile,
                               "stages_s": {"voxel_subsampling": sec["voxel_subsampling"], "median_resolution": sec["median_resolution"],
                                            "descriptor_nn": per_row * n_sub[0] + t_build, "pooling": sec["pooling"], "coarse": sec["coarse"],
                                            "fine_scaled": fine_full, "merge_of_sample": sec["merge"]},
                               "sample": "one tile, stage by stage (oracle/paths.py c2f_tile): descriptor search %d rows x %d (%s) scaled to "
                                         "%d rows; fine matching of %d of %d patch pairs (one Python process, like the reference's loop) "
                                         "scaled by source points; all other stages on the whole tile" %
                                         (sample_rows, n_sub[1], how, n_sub[0], pairs_done, pairs_all),
                               "backends_present": backends}
        # in-run parity of the sampled pairs against the GPU result of the same tile
        flips = bad_status = bad_K = checked = fitted = pair_lists_equal = 0
        worst = 0.0
        for lv, l in enumerate(o["levels"]):
            k = len(l["m"])
            fr = c.fine_results_multiple[lv]
            of = l["fine"]
            firsts_g = [int(x[0]) for x in c.data_output.spt_corres_src_multiple[lv][:k]]
            _, spt_s = opaths.patch_lists(tt["partition_src"][lv].cpu().numpy(), 10)
            if firsts_g != [int(spt_s[m_][0]) for m_ in l["m"]]:
                continue
            pair_lists_equal += 1
            Kg, Sg, Ig = (x[:k].cpu().numpy() for x in (fr.K, fr.status, fr.iters))
            checked += k
            bad_K += int((Kg != of["K"]).sum())
            bad_status += int((Sg != of["status"]).sum())
            ok = (of["status"] == 0) & (Sg == 0)
            fitted += int(ok.sum())
            flips += int((ok & (Ig != of["iters"])).sum())
            Tg = fr.T[:k].cpu().numpy().astype(np.float64)
            for q in np.nonzero(ok & (Ig == of["iters"]))[0]:
                P = tt["src_pts"][torch.as_tensor(spt_s[l["m"][q]])].cpu().numpy().astype(np.float64)
                To = of["T"][q].astype(np.float64)
                worst = max(worst, float(np.abs((P @ Tg[q][:3, :3].T + Tg[q][:3, 3]) - (P @ To[:3, :3].T + To[:3, 3])).max()))
        from fusion4landslide_b200 import ops as _ops
        di_ = c.data_interim
        _, _, desc_tie = _ops.desc_nn(di_.tile_pts_sub_feat_src, di_.tile_pts_sub_feat_tgt, tie_eps=1e-6)
        knn_tie = _ops.knn_ties(di_.src_pts_sub.contiguous(), c.data_input_3d.src_pts.contiguous(), 1)
        res["ties"] = {"desc_nn_rows_within_1e-6": int(desc_tie.sum()), "of_rows": int(desc_tie.numel()),
                       "voxel_to_point_nn_ties_rel_1e-6": int(knn_tie.sum()), "of_voxels": int(knn_tie.numel()),
                       "note": "flags exported by f4l_desc_nn_ex / f4l_knn_grid_ties on one tile: the rows exempt from index comparison"}
        res["parity"] = {"checked_against": "oracle/paths.py on the CPU-sampled pairs of one tile, in this run",
                         "levels_with_identical_pair_lists": pair_lists_equal, "pairs_checked": checked, "K_mismatch": bad_K,
                         "status_mismatch": bad_status, "fitted_pairs": fitted, "icp_path_flips": flips,
                         "max_abs_dvf_diff_m_same_path": worst,
                         "ties_2d_vote": int(sum(int(l["tie"].sum()) for l in o["levels"]))}
    del tiles
    torch.cuda.empty_cache()
    return res


def run_configs(a, dev, L):
    peaks = _peaks_full()
    want = [] if a.configs.lower() == "none" else [c.strip().upper() for c in a.configs.split(",") if c.strip()]
    out = {}
    fns = {"C1": lambda: bench_c1(a, dev, L, peaks), "C2": lambda: bench_c2(a, dev, L, peaks),
           "C3": lambda: bench_c2f(a, dev, L, peaks, False), "C4": lambda: bench_c2f(a, dev, L, peaks, True)}
    for name in want:
        t0 = time.perf_counter()
        try:
            out[name] = fns[name]()
        except Exception as e:                       # one config failing must not take the headline line down
            import traceback
            out[name] = {"error": "%s: %s" % (type(e).__name__, e), "traceback": traceback.format_exc()[-1500:]}
            torch.cuda.empty_cache()
        out[name]["wall_s"] = round(time.perf_counter() - t0, 1)
    return out


def run_dips(a):
    """--workload dips: DIPs patch front-end (src/data_loader.py:16-109) at tile scale, single GPU.
    One step = index build + (n,3,256) patches of every point of one C3-shaped tile-epoch (625 k points at 0.1 m
    spacing, feature radius sqrt(3)*10*resolution, ~850 neighbours per point).  cpu_baseline: oracle/dips.py (numpy +
    cKDTree, one core -- the reference's own per-point loop shape) on a bounded sample of the same queries."""
    if int(os.environ.get("RANK", "0")) != 0:
        return                                      # single-GPU workload: under torchrun only rank 0 runs it
    import numpy as np
    import torch
    from fusion4landslide_b200 import _lib, ops, synth
    dev = torch.device("cuda:0")
    peak, peak_src = load_peaks()
    n = a.dips_pts
    batch = 125_000
    d = synth.make_tile(n, seed=3, device=dev)
    ref = d["src"].double().contiguous()
    radius = float(np.sqrt(3) * 10 * 0.1)
    L = _lib.lib()
    out = torch.empty((batch, 3, 256), dtype=torch.float32, device=dev)

    def step():
        index = ops.DipsIndex(ref, radius)
        cnt = []
        for off in range(0, n, batch):
            q = ref[off:off + batch]
            cnt.append(ops.dips_patches(index, q, 256, seed=off, out=out[:q.shape[0]])[1])
        return torch.cat(cnt)

    for _ in range(max(a.warmup, 1)):
        cnt = step()
    torch.cuda.synchronize()
    L.f4l_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(0)
    sampler.start()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = int(L.f4l_launch_count())
    ms = e0.elapsed_time(e1) / a.steps
    L.f4l_profile_reset(); L.f4l_profile_enable(1)
    step()
    torch.cuda.synchronize()
    L.f4l_profile_enable(0)
    kern = {k: {"ms_total": round(v[0], 4), "launches": v[1]} for k, v in _lib.profile_table().items()}
    kms = kern.get("k_dips_patches", {}).get("ms_total", ms)
    alg = n * (24 + 3 * 256 * 4)
    line = {"metric": "dips_patches_per_sec", "value": n / (ms * 1e-3), "unit": "patches/s", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (f32 out)", "data": "synthetic",
            "config": {"workload": "DIPs front-end: 1 tile-epoch of %d points (cloud = queries), radius %.3f m, 256-point patches"
                                   % (n, radius), "neighbours_mean": float(cnt.float().mean()), "neighbours_max": int(cnt.max()),
                       "l2": "3 KB written per point (1.9 GB per step) exceeds the L2"},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"kernel": "k_dips_patches", "bound": "hbm", "achieved": alg / (kms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (kms * 1e-3) / 1e9 / peak, "traffic": (ncu_traffic("k_dips_patches") or [None])[0],
                         "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                         "note": "24 B in + 3*256*4 B out per point; the kernel is FP64 / latency-bound today"},
            "kernels": kern}
    if not a.no_cpu_baseline:
        import time
        from scipy.spatial import cKDTree
        from oracle import dips as odips
        refn = ref.cpu().numpy()
        rng = np.random.default_rng(0)
        pick = rng.choice(n, a.dips_cpu_queries, replace=False)
        tree = cKDTree(refn)
        t0 = time.perf_counter()
        for q in refn[pick]:
            pa, _, _ = odips.extract_all(q, tree, refn, radius)
            odips.sample(pa, rng.permutation(max(pa.shape[0], 256))[:256])
        t_q = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": a.dips_cpu_queries / t_q, "unit": "patches/s", "cores": 1, "kind": "port",
                                "sample": "%d queries of the same tile (oracle/dips.py: numpy + cKDTree, tree build excluded)" % a.dips_cpu_queries}
    print(json.dumps(line))


def run_reference(a):
    """--impl reference: the CPU arm alone, on this box's host cores.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from fusion4landslide_b200 import pipeline, synth
    cfg = pipeline.FineConfig()
    d = make_c5_tile(a, a.seed * 100003, "cpu")
    tile = pipeline.prepare_tile(d["src"], d["tgt"], d["label_src"], d["label_tgt"], d["corr3d"])
    workers = os.cpu_count() or 1
    per_step = a.cpu_sample_pairs or max(32, min(tile.n_pairs, int(8.0 * workers * 150)))
    vals, secs = [], []
    for s in range(a.warmup + a.steps):
        r = cpu_arm_sample(a, tile, cfg, workers=workers, pairs=per_step)
        if s >= a.warmup:
            vals.append(r["value"])
            secs.append(r["seconds"])
        if s == 0 and r["seconds"] > 40:        # keep the whole run within a few minutes
            per_step = max(32, int(per_step * 20 / r["seconds"]))
    v = sum(vals) / len(vals)
    ms = 1e3 * sum(secs) / len(secs)
    r["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (numpy/scipy), f32 i/o", "data": "synthetic",
            "config": {"workload": workload_name(a), "tiles": a.tiles, "tile_pts": a.tile_pts, "scene": a.scene,
                       "icp_threshold": cfg.icp_threshold, "assign_type": cfg.assign_type,
                       "step": "bounded sample of the same workload: %d patch pairs of one tile per step, all host cores" % per_step},
            "cpu_baseline": r,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "dips":
        run_dips(args)
    else:
        run_b200(args)
