#!/usr/bin/env python3
"""Top source lines of a kernel from an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_hot_lines.py report.ncu-rep kernel_name [n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True).stdout.decode()
lines = []
for r in csv.reader(out.splitlines()):
    if len(r) > 8 and r[0].isdigit() and r[7].isdigit():
        samp = int(r[4]) if r[4].isdigit() else 0
        lines.append((int(r[7]), samp, int(r[0]), r[1].strip()[:120]))
tot = sum(l[0] for l in lines) or 1
ts = sum(l[1] for l in lines) or 1
print("total warp instructions %d, stall samples %d" % (tot, ts))
print("--- by instructions executed")
for inst, samp, ln, src in sorted(lines, reverse=True)[:n]:
    print("%5.1f%% inst %5.1f%% samp  L%-4d %s" % (100 * inst / tot, 100 * samp / ts, ln, src))
print("--- by stall samples")
for inst, samp, ln, src in sorted(lines, key=lambda x: -x[1])[:n]:
    print("%5.1f%% inst %5.1f%% samp  L%-4d %s" % (100 * inst / tot, 100 * samp / ts, ln, src))
