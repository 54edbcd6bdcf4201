#!/usr/bin/env python3
"""Host-thread scaling of the pipelined search on a multi-GPU box: usage search_threads_check.py <devices>"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import gpu_ai_b200 as b  # noqa: E402

START = np.array([0x00000FFF, 0xFFF00000, 0, 0], dtype=np.uint32)
devices = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = b.Engine(devices=devices)
print(json.dumps({"host_cpus": len(os.sched_getaffinity(0)), "devices": devices}), flush=True)
for reps, threads in ((32, 16), (32, 32), (32, 0), (64, 0), (32, 48)):
    t = b.Tree(START)
    t.search_ex(eng, iterations=3, initial_batch=65536, max_batch=1 << 20, reps=reps, depth=2, threads=threads)
    t = b.Tree(START)
    st = t.search_ex(eng, seconds=1.0, initial_batch=65536, scale=0.02, max_batch=1 << 20, reps=reps, key=3, depth=2, threads=threads)
    print(json.dumps({"reps": reps, "threads_requested": threads, "threads": st["threads"], "playouts_per_s": st["playouts"] / st["seconds"],
                      "leaf_selections_per_s": st["leaves"] / st["seconds"], "gpu_busy": st["kernel_s"] / st["seconds"],
                      "select_s": st["select_s"], "update_s": st["update_s"], "wait_s": st["wait_s"]}), flush=True)
