#!/usr/bin/env python
"""Top CUDA source lines of a kernel by warp-stall samples, from an ncu report captured with
--import-source on (needs -lineinfo).   python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, lines = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        try:
            lines.append((int(r[6]), int(r[7]), cur, r[0], r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
print("total samples", tot, " total warp instr", sum(l[1] for l in lines))
for n, ex, f, ln, src in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% %9d  %s:%s  %s" % (100.0 * n / tot, ex, f, ln, src))
