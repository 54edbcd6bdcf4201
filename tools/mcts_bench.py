#!/usr/bin/env python3
"""Search throughput of b2p_tree_search (tree on the host, playouts on the GPU) from the initial position:
playouts/s and tree size for several (batch, reps) policies at a fixed wall-clock budget.  The reference's
mcts_host reaches ~7e4 trials per 7 s move."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_ai_b200 as b  # noqa: E402
from oracle.pyoracle import START_PACKED  # noqa: E402

eng = b.Engine(devices=1)
for batch, scale, reps in ((50, 0.0, 1), (4000, 0.0, 1), (4000, 0.02, 1), (4000, 0.02, 16), (16384, 0.02, 64), (65536, 0.02, 256)):
    t = b.Tree(START_PACKED)
    t.search(eng, iterations=2, initial_batch=batch, scale=scale, reps=reps)   # warm-up
    t = b.Tree(START_PACKED)
    t0 = time.perf_counter()
    played = t.search(eng, seconds=1.0, initial_batch=batch, scale=scale, reps=reps, key=3)
    dt = time.perf_counter() - t0
    info = t.info()
    print(json.dumps({"initial_batch": batch, "scale": scale, "reps": reps, "seconds": round(dt, 3), "playouts": played,
                      "playouts_per_s": played / dt, "tree_nodes": info["nodes"], "leaf_selections_per_s": played / reps / dt}), flush=True)
