#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --set full` report of `bench.py` (dominant kernel, one launch):
dram__bytes_read.sum + dram__bytes_write.sum and the headline counters.  Usage:
    python tools/extract_traffic.py <report.ncu-rep> <leaves> <reps> <mode> <order>"""
import csv
import json
import os
import subprocess
import sys

rep, leaves, reps, mode, order = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def bytes_of(name):
    u, v = m[name]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return float(v) * scale


keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
res = {"report": os.path.basename(rep), "kernel": m["Kernel Name"][1] if "Kernel Name" in m else None, "leaves": leaves, "reps": reps,
       "mode": mode, "order": order, "dram_bytes_read": bytes_of("dram__bytes_read.sum"), "dram_bytes_write": bytes_of("dram__bytes_write.sum")}
res["dram_bytes_per_launch"] = res["dram_bytes_read"] + res["dram_bytes_write"]
res["algorithmic_bytes_per_launch"] = 16 * leaves + leaves * reps
res["counters"] = {k: (float(m[k][1]) if k in m else None) for k in keep}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
json.dump(res, open(path, "w"), indent=1)
print(json.dumps(res, indent=1))
