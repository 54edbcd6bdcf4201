#!/usr/bin/env python3
"""Host-side throughput of the search tree (no GPU): leaf selections/s of b2p_tree_select_batch + b2p_tree_update_batch
(the two host halves of a pipelined search round) and of the serial interface, from the initial position, with fake
playout results.  Usage: tree_bench.py [threads ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_ai_b200 as b  # noqa: E402

START = np.array([0x00000FFF, 0xFFF00000, 0, 0], dtype=np.uint32)
threads_list = [int(x) for x in sys.argv[1:]] or [1, 4, os.cpu_count() or 1]
rng = np.random.default_rng(1)
for batch in (4096, 65536, 1 << 20):
    wins_pool = rng.integers(0, 5, size=(batch, 2)).astype(np.uint32)
    rounds = max(4, min(200, (1 << 23) // batch))
    t = b.Tree(START)
    t0 = time.perf_counter()
    sel = upd = 0.0
    for it in range(rounds):
        a = time.perf_counter()
        t.select(batch)
        c = time.perf_counter()
        t.update_counts(wins_pool, 8)
        sel += c - a
        upd += time.perf_counter() - c
    print(json.dumps({"interface": "serial b2p_tree_select/update_counts", "batch": batch, "rounds": rounds,
                      "leaves_per_s": batch * rounds / (sel + upd), "select_s": sel, "update_s": upd, "nodes": t.info()["nodes"]}), flush=True)
    for th, exact in [(x, e) for x in threads_list for e in (True, False)]:
        t = b.Tree(START)
        sel = upd = 0.0
        for it in range(rounds):
            a = time.perf_counter()
            t.select_batch(0, batch, reps=8, threads=th, exact=exact)
            c = time.perf_counter()
            t.update_batch(0, wins_pool, threads=th)
            sel += c - a
            upd += time.perf_counter() - c
        print(json.dumps({"interface": "b2p_tree_select_batch/update_batch", "threads": th, "exact": exact, "batch": batch, "rounds": rounds,
                          "leaves_per_s": batch * rounds / (sel + upd), "select_s": sel, "update_s": upd, "nodes": t.info()["nodes"]}), flush=True)
