#!/usr/bin/env python3
"""Integer-pipe issue rates of this GPU (b2p_microbench): freezes the INT32 roofline denominator."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_ai_b200 as b  # noqa: E402

NAMES = ["LOP3", "IADD3", "SHF", "POPC", "IMAD", "LOP3+IMAD 1:1", "BREV", "FLO(bfind)", "LOP3+IMAD 3:1", "IMAD.HI (mul.hi)", "LOP3 + mul.hi-as-shift 1:1", "LOP3 + SHF.R 1:1"]
eng = b.Engine(devices=1)
info = eng.device_info(0)
res = {"gpu": info["name"], "sms": info["sm_count"], "clock_khz_max": info["clock_khz"], "rates": {}}
for which, name in enumerate(NAMES):
    best = 0.0
    for _ in range(3):
        ops, ms = eng.microbench(which, iters=4000)
        best = max(best, ops)
    per_sm_clk = best / (info["sm_count"] * info["clock_khz"] * 1e3)
    res["rates"][name] = {"thread_ops_per_s": best, "per_sm_per_clk_at_max_clock": per_sm_clk, "ms": ms}
print(json.dumps(res, indent=1))
