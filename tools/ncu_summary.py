#!/usr/bin/env python
"""Turn the scratch ncu outputs under gpurun_out/ into the small tracked summaries under profiles/.

    python tools/ncu_summary.py r1            # prefix of the files written

Reads gpurun_out/prof_<kernel>.ncu-rep (ncu --set full captures) and gpurun_out/launches.csv
(the --metrics gpu__time_duration.sum launch list).  Needs the `ncu` CLI (no GPU).
"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SCRATCH = os.path.join(ROOT, "gpurun_out")

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__maximum_warps_per_active_cycle_pct",
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarise_rep(rep, prefix):
    name = os.path.basename(rep)[len("prof_"):-len(".ncu-rep")]
    hdr, units, rows = raw_page(rep)
    lines = ["# ncu --set full --clock-control none : %s  (from gpurun_out/%s)" % (name, os.path.basename(rep)), ""]
    for r in rows:
        kn = r[hdr.index("Kernel Name")]
        lines.append("kernel: %s" % kn[:160])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append("  %-78s %s %s" % (m, r[i], units[i]))
        lines.append("")
    path = os.path.join(OUT, "%s_ncu_%s.txt" % (prefix, name))
    open(path, "w").write("\n".join(lines))
    return path


def summarise_launches(path, prefix):
    rows = list(csv.reader(open(path)))
    start = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    agg = collections.OrderedDict()
    order = []
    for r in rows[start + 1:]:
        n = r[4].split("(")[0].replace("void ", "")
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
        order.append((n, float(r[-1]), r[7], r[8]))
    tot = sum(v[1] for v in agg.values()) or 1.0
    out = ["# launch list summary (ncu --metrics gpu__time_duration.sum --clock-control none); cold-cache, serialised:",
           "# compare SHARES with bench.py's `kernels` table, not absolutes", "kernel,launches,total_us,share"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s,%d,%.1f,%.4f" % (n[-90:], c, t / 1e3, t / tot))
    p = os.path.join(OUT, "%s_launches_summary.csv" % prefix)
    open(p, "w").write("\n".join(out) + "\n")
    full = os.path.join(OUT, "%s_launches.csv" % prefix)
    with open(full, "w") as f:
        f.write("kernel,duration_ns,block,grid\n")
        for n, t, b, g in order:
            f.write('"%s",%d,"%s","%s"\n' % (n[-90:], t, b, g))
    return p


if __name__ == "__main__":
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    for rep in sorted(glob.glob(os.path.join(SCRATCH, "prof_*.ncu-rep"))):
        print(summarise_rep(rep, prefix))
    lc = os.path.join(SCRATCH, "launches.csv")
    if os.path.exists(lc):
        print(summarise_launches(lc, prefix))
