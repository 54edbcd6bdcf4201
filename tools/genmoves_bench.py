#!/usr/bin/env python3
"""Throughput of the batched move-list kernel (replaces genMovesKernel/genMovesTest, src/genMovesTest.cu:10-100):
device-resident D_ref leaves, CUDA events, median of 10.  HBM bytes = 16 B in + 1 B count + 8 B per emitted move."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_ai_b200 as b  # noqa: E402

eng = b.Engine(devices=[0])
dev = torch.device("cuda", 0)
N, MAXM = 1 << 20, 32
s = torch.empty((N, 4), dtype=torch.int32, device=dev)
eng.gen_leaves_device(N, s.data_ptr(), key=2016)
moves = torch.zeros((N, MAXM), dtype=torch.int64, device=dev)
counts = torch.zeros(N, dtype=torch.uint8, device=dev)
ms = []
for it in range(13):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.genmoves_device(s.data_ptr(), N, MAXM, moves.data_ptr(), counts.data_ptr(), stream=0)
    c.record()
    torch.cuda.synchronize()
    if it >= 3:
        ms.append(a.elapsed_time(c))
t = float(np.median(ms)) * 1e-3
emitted = int(counts.sum().item())
bytes_alg = 17 * N + 8 * emitted
print(json.dumps({"states": N, "max_moves": MAXM, "ms": t * 1e3, "states_per_s": N / t, "moves_per_s": emitted / t,
                  "algorithmic_GBs": bytes_alg / t / 1e9, "hbm_peak_GBs": 6539.2, "frac": bytes_alg / t / 1e9 / 6539.2,
                  "note": "the reference tests one state per launch with 3 cudaMallocs (src/genMovesTest.cu:26-100)"}))
