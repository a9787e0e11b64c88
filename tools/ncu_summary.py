#!/usr/bin/env python3
"""Text summary of an .ncu-rep (key raw metrics per profiled launch + per-function / per-line shares).
    python tools/ncu_summary.py report.ncu-rep [source.cu] [--json profiles/issue_profile.json] > profiles/<name>.txt
--json also writes the per-kernel pipe utilisation / issue figures bench.py quotes in roofline.issue."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]

json_out = None
if "--json" in sys.argv:
    i = sys.argv.index("--json")
    json_out = sys.argv[i + 1]
    del sys.argv[i:i + 2]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True).stdout.decode()
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print("# ncu --set full summary of %s" % rep.split("/")[-1])
kernels = []
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    kernels.append(name)
    print("\n== launch id %s: %s" % (r[idx["ID"]], name))
    for k in KEYS:
        if k in idx:
            print("  %-82s %s %s" % (k, r[idx[k]], units[idx[k]]))
if json_out:
    import json
    want = {"time_ms": "gpu__time_duration.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "pipe_xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "pipe_fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "pipe_lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum",
            "l2_hit_pct": "lts__t_sector_hit_rate.pct", "warp_instructions": "smsp__inst_executed.sum",
            "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "l2_atomic_sectors": "lts__t_sectors_op_atom.sum", "l2_red_sectors": "lts__t_sectors_op_red.sum"}
    prof = {"source": rep.split("/")[-1], "kernels": {}}
    for r in rows[2:]:
        short = r[idx["Kernel Name"]].split("(")[0].split("::")[-1].replace("void ", "").strip()
        if short in prof["kernels"]:
            continue
        d = {}
        for k, m in want.items():
            if m in idx:
                try:
                    d[k] = float(r[idx[m]].replace(",", ""))
                except ValueError:
                    pass
                d.setdefault("units", {})[k] = units[idx[m]]
        prof["kernels"][short] = d
    prof["popc_floor"] = ("screen: 56 (window, period) tests per 150-base read x 6 POPC at 16 lanes/clk/SM = 21 clk per read per SM "
                          "-> 13.8 G reads/s = 0.83 TB/s = 12.7 % of the measured HBM peak (DESIGN.md section 4)")
    with open(json_out, "w") as f:
        json.dump(prof, f, indent=1)
        f.write("\n")
if len(sys.argv) > 2:
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    seen = set()
    for name in kernels:
        short = name.split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
        if short in seen:
            continue
        seen.add(short)
        print("\n== %s: share of warp instructions / stall samples per source function" % short)
        print(subprocess.run([sys.executable, os.path.join(here, "ncu_by_function.py"), rep, short, sys.argv[2]],
                             capture_output=True).stdout.decode().rstrip())
        print("\n== %s: hottest source lines" % short)
        print(subprocess.run([sys.executable, os.path.join(here, "ncu_hot_lines.py"), rep, short, "15"],
                             capture_output=True).stdout.decode().rstrip())
