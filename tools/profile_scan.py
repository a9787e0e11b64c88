#!/usr/bin/env python3
"""Small driver for ncu: one device-resident batch of the configs[1] shape, scanned a few times.

    python tools/profile_scan.py [reads] [scans] [min_mer] [max_mer] [tel_ppm] [half_ppm] [n_ppm] [mode 0 short / 1 pair / 2 long] [read_len] [flavor]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trew_b200 import api  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
scans = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mn = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mx = int(sys.argv[4]) if len(sys.argv) > 4 else 32
tel = int(sys.argv[5]) if len(sys.argv) > 5 else 10000
half = int(sys.argv[6]) if len(sys.argv) > 6 else 2000
nppm = int(sys.argv[7]) if len(sys.argv) > 7 else 1000
mode = int(sys.argv[8]) if len(sys.argv) > 8 else api.MODE_SHORT
read_len = int(sys.argv[9]) if len(sys.argv) > 9 else 150
flavor = int(sys.argv[10]) if len(sys.argv) > 10 else 0
with api.DeviceContext(mode, mn, mx) as ctx:
    h = ctx.synth_resident(1, reads, read_len, tel_ppm=tel, half_ppm=half, n_ppm=nppm, sub_ppm=1000 if flavor == 2 else 10000, flavor=flavor)
    for _ in range(2):          # warm-up: first launches load the kernels (lazy module loading costs milliseconds)
        ctx.scan_resident(h)
    ctx.sync()
    ctx.kernel_times()
    for _ in range(scans):
        ctx.scan_resident(h)
    ctx.sync()
    s, d, e, n = ctx.kernel_times()
    st = ctx.stats()
    print("reads %d scans %d screen %.3f decide %.3f exact %.3f ms/scan survivors %.4f" %
          (reads, n, s / n, d / n, e / n, st.survivors / st.units))
    ctx.free_resident(h)
