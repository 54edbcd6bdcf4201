#!/usr/bin/env python3
"""The reference's benchmark protocol (tests.sh:10-36: `run_ai -m playout_test -n N -1 D -2 D` for
N in 50..200000, 5 repeats) through the B200 drop-in drivers, as JSON lines instead of appended stdout.

For every driver name and batch size: wall-clock of `runPlayouts(states)` on N reference `State`
objects (776 B each, host memory) exactly as playoutTest times it (src/driver.cpp:119-123), best and
median of 5, plus the win tallies it prints.  Leaves = D_ref (reproducible genRandomStates)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_ai_b200 as b  # noqa: E402
from gpu_ai_b200 import engine as E  # noqa: E402

SIZES = [50, 100] + list(range(200, 1001, 200)) + list(range(2000, 10001, 2000)) + list(range(20000, 100001, 20000)) + [200000]
devices = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = b.Engine(devices=devices)
leaves = eng.gen_leaves(max(SIZES), key=2016)
s776 = E.unpack776(leaves)
for name in ("device_single", "device_multiple", "device_coarse", "device_heuristic"):
    drv = b.getPlayoutDriver(name)
    drv.engine = eng
    for n in SIZES:
        times = []
        for rep in range(6):
            t0 = time.perf_counter()
            res = drv.runPlayouts(s776[:n])
            dt = time.perf_counter() - t0
            if rep:
                times.append(dt)
        print(json.dumps({"driver": name, "gpus": devices, "n": n, "best_s": min(times), "median_s": float(np.median(times)),
                          "playouts_per_s": n / min(times), "draws": int((res == -1).sum()), "p1": int((res == 0).sum()),
                          "p2": int((res == 1).sum())}), flush=True)
