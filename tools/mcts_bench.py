#!/usr/bin/env python3
"""Search throughput of b2p_tree_search_ex (tree on the host, playouts on the GPU(s)) from the initial position.
Usage: mcts_bench.py [devices] [seconds].  One JSON line per (batch policy, reps, depth, threads)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_ai_b200 as b  # noqa: E402
import numpy as np  # noqa: E402

START = np.array([0x00000FFF, 0xFFF00000, 0, 0], dtype=np.uint32)
devices = int(sys.argv[1]) if len(sys.argv) > 1 else 1
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
eng = b.Engine(devices=devices)
for batch, scale, mx, reps, depth, threads, policy in [x + (0,) for x in (
        (50, 0.0, 50, 1, 1, 1), (4096, 0.0, 4096, 1, 2, 0), (4096, 0.0, 4096, 32, 2, 0), (65536, 0.02, 1 << 20, 1, 2, 0),
        (65536, 0.02, 1 << 20, 8, 2, 0), (65536, 0.02, 1 << 20, 16, 2, 0), (65536, 0.02, 1 << 20, 32, 1, 0),
        (65536, 0.02, 1 << 20, 32, 2, 0), (65536, 0.02, 1 << 20, 32, 3, 0), (262144, 0.0, 262144, 32, 2, 0),
        (1 << 20, 0.0, 1 << 20, 32, 2, 0), (65536, 0.02, 1 << 20, 64, 2, 0), (65536, 0.02, 1 << 20, 256, 2, 0))] + [
        (65536, 0.02, 1 << 20, 32, 2, 0, 1), (2048, 0.0, 2048, 16, 2, 0, 1), (8192, 0.0, 8192, 8, 2, 0, 1)]:
    t = b.Tree(START)
    t.search_ex(eng, iterations=3, initial_batch=batch, scale=scale, max_batch=mx, reps=reps, depth=depth, threads=threads, policy=policy)   # warm-up
    t = b.Tree(START)
    st = t.search_ex(eng, seconds=seconds, initial_batch=batch, scale=scale, max_batch=mx, reps=reps, key=3, depth=depth, threads=threads, policy=policy)
    st.update({"devices": devices, "initial_batch": batch, "scale": scale, "max_batch": mx, "reps": reps, "policy": "uct" if policy else "reference",
               "playouts_per_s": st["playouts"] / st["seconds"], "leaf_selections_per_s": st["leaves"] / st["seconds"],
               "gpu_busy": st["kernel_s"] / st["seconds"],
               "other_s": st["seconds"] - st["select_s"] - st["update_s"] - st["wait_s"]})
    print(json.dumps(st), flush=True)
