#!/usr/bin/env python3
"""Thread-per-playout vs warp-per-playout: kernel latency and throughput over batch size (device-resident
leaves, CUDA events, median of 20).  Feeds the B2P_SCHED_AUTO threshold and DESIGN.md 4.5."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_ai_b200 as b  # noqa: E402

eng = b.Engine(devices=[0])
dev = torch.device("cuda", 0)
N = 1 << 20
d_states = torch.empty((N, 4), dtype=torch.int32, device=dev)
eng.gen_leaves_device(N, d_states.data_ptr(), key=2016)
d_w = torch.empty(N, dtype=torch.int8, device=dev)
d_c = torch.zeros(4, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
rows = []
for mode, mname in ((b.MODE_RANDOM, "random"), (b.MODE_HEURISTIC, "heuristic")):
    order = b.ORDER_FAST if mode == b.MODE_RANDOM else b.ORDER_CANONICAL
    for n in (32, 50, 128, 512, 2048, 4096, 8192, 32768, 131072, 1 << 20):
        rec = {"mode": mname, "n": n}
        for sched, sname in ((b.SCHED_THREAD, "thread"), (b.SCHED_WARP, "warp")):
            if sname == "warp" and n > 131072 and mode == b.MODE_HEURISTIC:
                continue
            ms = []
            for it in range(23):
                a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                eng.run_packed_device(d_states.data_ptr(), n, reps=1, key=100 + it, mode=mode, sched=sched, order=order,
                                      d_winners=d_w.data_ptr(), d_counters=d_c.data_ptr(), stream=0)
                bb.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ms.append(a.elapsed_time(bb))
            rec[sname + "_us"] = round(1e3 * float(np.median(ms)), 1)
            rec[sname + "_playouts_per_s"] = n / (float(np.median(ms)) * 1e-3)
        rows.append(rec)
        print(json.dumps(rec), flush=True)
