#!/usr/bin/env python3
"""Summarise an ncu report's SASS page into basic blocks: executions, share of warp instructions, active lanes.
Usage: python tools/ncu_blocks.py <report.ncu-rep> [min_share_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ie, it = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
data = [(r[ia].strip(), int(r[ie]), float(r[it])) for r in rows[2:] if len(r) > it]
tot = sum(d[1] for d in data)
cands = [d[1] for d in data if d[2] >= 31.5 and "WARPSYNC" in d[0]]
loop = max(cands) if cands else 0
print("kernel:", rows[0][1][:90])
print("total warp instructions %d; ply-loop iterations %d; warp-instr per iteration %.0f; lane-weighted avg threads %.2f"
      % (tot, loop, tot / loop if loop else 0, sum(d[1] * d[2] for d in data) / tot))
i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and data[j + 1][1] == data[i][1]:
        j += 1
    n, cnt = j - i + 1, data[i][1]
    thr = sum(d[2] for d in data[i:j + 1]) / n
    ops = " ".join((d[0].split()[1] if d[0].startswith("@") else d[0].split()[0]) for d in data[i:min(j + 1, i + 7)])
    if 100.0 * n * cnt / tot >= min_share:
        print("%4d-%4d n=%3d exec=%10d share=%5.1f%% lanes=%5.1f | %s" % (i, j, n, cnt, 100.0 * n * cnt / tot, thr, ops))
    i = j + 1
