#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an .ncu-rep (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = ""
hdr = None
lines = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = hdr.index("Instructions Executed")
        cs = hdr.index("# Samples")
        continue
    if hdr and r[0] and r[0].isdigit() and len(r) > ci:
        try:
            lines[(cur_file, int(r[0]))] = (int(r[ci]), int(r[cs] or 0), r[1].strip())
        except ValueError:
            pass
tot = sum(v[0] for v in lines.values())
ts = sum(v[1] for v in lines.values())
print("total warp instructions %d, stall samples %d" % (tot, ts))
for (f, ln), (n, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d | %s" % (100 * s / max(ts, 1), 100 * n / max(tot, 1), f, ln, src[:100]))
