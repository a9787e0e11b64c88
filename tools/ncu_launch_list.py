#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/ncu_launch_list.py launches.csv > profiles/<name>.txt"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit = r.get("Metric Unit", "ns")
        v = float(r["Metric Value"].replace(",", ""))
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        name = r["Kernel Name"].split("(")[0]
        rows.append((name, ms))
    tot = sum(ms for _, ms in rows)
    agg = OrderedDict()
    for name, ms in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    print("# %d launches, %.3f ms in total" % (len(rows), tot))
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s launches %4d  total %10.3f ms  share %5.1f%%  mean %8.3f ms" % (name[:60], n, ms, 100 * ms / tot, ms / n))
    scan = {k: v[1] for k, v in agg.items() if "trew_screen" in k or "trew_filter" in k or "trew_exact" in k}
    if scan:
        st = sum(scan.values())
        print("# share among the three scan kernels only:")
        for k, v in sorted(scan.items(), key=lambda kv: -kv[1]):
            print("#   %-60s %5.1f%%" % (k[:60], 100 * v / st))


if __name__ == "__main__":
    main()
