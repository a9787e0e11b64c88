#!/usr/bin/env python3
"""Instruction / stall-sample share per source function of a kernel from an .ncu-rep.
    python tools/ncu_by_function.py report.ncu-rep kernel_name source.cu"""
import csv
import re
import subprocess
import sys

rep, kern, src = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True).stdout.decode()
starts = []
for i, line in enumerate(open(src), 1):
    m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|__host__)[^;]*?\b(\w+)\s*\(", line)
    if m and not line.startswith(" "):
        starts.append((i, m.group(1)))
def fn_of(ln):
    name = "?"
    for s, n in starts:
        if s <= ln:
            name = n
        else:
            break
    return name
inst, samp = {}, {}
cur_file = None
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1]
    if len(r) > 8 and r[0].isdigit() and r[7].isdigit():
        f = fn_of(int(r[0])) if cur_file and cur_file.endswith(src.split("/")[-1]) else "<" + (cur_file or "?").split("/")[-1] + ">"
        inst[f] = inst.get(f, 0) + int(r[7])
        samp[f] = samp.get(f, 0) + (int(r[4]) if r[4].isdigit() else 0)
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
for f in sorted(inst, key=lambda k: -inst[k]):
    print("%5.1f%% inst %5.1f%% samp  %s" % (100 * inst[f] / ti, 100 * samp[f] / ts, f))
