#!/usr/bin/env python3
"""Per-source-line stall breakdown of a kernel from an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_stalls_by_line.py report.ncu-rep kernel_name [stall_column] [n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
col = sys.argv[3] if len(sys.argv) > 3 else "stall_long_sb"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True).stdout.decode()
hdr, agg = None, {}
for r in csv.reader(out.splitlines()):
    if len(r) > 1 and r[1] == "Source":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        def g(k):
            try:
                return int(d.get(k, "0") or 0)
            except ValueError:
                return 0
        agg[int(r[0])] = (d["Source"].strip()[:100], g(col), g("# Samples"), g("Instructions Executed"))
tot = sum(v[1] for v in agg.values()) or 1
ts = sum(v[2] for v in agg.values()) or 1
print("%s total %d of %d samples" % (col, tot, ts))
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
    print("L%-5d %6d (%4.1f%%)  samples %6d  inst %9d  %s" % (ln, v[1], 100.0 * v[1] / tot, v[2], v[3], v[0]))
