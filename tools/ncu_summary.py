#!/usr/bin/env python3
"""Summarise .ncu-rep captures (ncu --set full) into profiles/: one text block per kernel with the
metrics DESIGN.md / bench.py quote, and profiles/ncu_traffic_r02.json (DRAM bytes per launch; BG_TRAFFIC_FILE overrides).

  python tools/ncu_summary.py OUT.txt REPORT.ncu-rep:CELLS[:NOTE] ...
"""
import csv
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    out = sys.argv[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tpath = os.path.join(root, "profiles", os.environ.get("BG_TRAFFIC_FILE", "ncu_traffic_r02.json"))
    traffic = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from `ncu --set full --clock-control none`; "
                       "cells = DP cells that launch processed (bench.py scales bytes/cell to its own launches)", "kernels": []}
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    text = []
    for spec in sys.argv[2:]:
        parts = spec.split(":")
        rep, cells, note = parts[0], float(parts[1]), (parts[2] if len(parts) > 2 else "")
        # REPORT may also be the `ncu -i X.ncu-rep --page raw --csv` text itself (made on the GPU box: the reports are
        # too large to travel back)
        raw = open(rep).read() if rep.endswith(".csv") else \
            subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
            name = d["Kernel Name"]
            text.append("== %s   [%s; %s]" % (name, os.path.basename(rep), note))
            for k in KEYS:
                if k in d:
                    text.append("   %-82s %-10s %s" % (k, u[k], d[k]))
            rd = float(d["dram__bytes_read.sum"]) * TO_BYTES[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * TO_BYTES[u["dram__bytes_write.sum"]]
            short = name.replace("void ", "").replace("bg::", "").split("(")[0]
            traffic["kernels"] = [e for e in traffic["kernels"] if e["kernel"] != short]
            traffic["kernels"].append({"kernel": short, "report": os.path.basename(rep), "note": note, "cells": cells,
                                       "dram_bytes_read": rd, "dram_bytes_write": wr,
                                       "time_ms": float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "s": 1e3}.get(u["gpu__time_duration.sum"], 1)})
            text.append("")
    open(out, "w").write("\n".join(text) + "\n")
    json.dump(traffic, open(tpath, "w"), indent=1)
    print("\n".join(text))


if __name__ == "__main__":
    main()
