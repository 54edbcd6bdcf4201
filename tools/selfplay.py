#!/usr/bin/env python3
"""Search-policy tuning for the B200 tree player: each candidate policy (batch size, growth, playouts per leaf, pipeline
depth) plays G games at a fixed wall-clock budget per move against the same opponent -- an mcts_host-equivalent
searcher (batches of 50, one playout per leaf, a fixed number of trials per move; same tree policy as the reference's
GameTree).  Usage: selfplay.py [games] [seconds per move] [opponent trials per move]"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_ai_b200 as b  # noqa: E402

START = np.array([0x00000FFF, 0xFFF00000, 0, 0], dtype=np.uint32)
games = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
opp_trials = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
eng = b.Engine(devices=1)
CANDIDATES = {
    "uct_b2048_r16": dict(initial_batch=2048, scale=0.0, max_batch=2048, reps=16, depth=2, policy=1),
    "uct_b8192_r8": dict(initial_batch=8192, scale=0.0, max_batch=8192, reps=8, depth=2, policy=1),
    "uct_b16384_grow_2^18_r16": dict(initial_batch=16384, scale=0.02, max_batch=1 << 18, reps=16, depth=2, policy=1),
    "uct_b512_r4": dict(initial_batch=512, scale=0.0, max_batch=512, reps=4, depth=2, policy=1),
    "ref_b2048_r16": dict(initial_batch=2048, scale=0.0, max_batch=2048, reps=16, depth=2, policy=0),
}


def wilson(p, n, z=1.96):
    den = 1 + z * z / n
    c = (p + z * z / (2 * n)) / den
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / den
    return [c - h, c + h]


for name, cfg in CANDIDATES.items():
    score = {"cand": 0, "opp": 0, "draw": 0}
    po, leaves = [], []
    t_start = time.time()
    for g in range(games):
        seat = ("cand", "opp") if g % 2 == 0 else ("opp", "cand")
        trees = [b.Tree(START), b.Tree(START)]
        plies = 0
        while True:
            info = trees[0].info()
            st = info["root_state"]
            turn, msc = int(st[3] & 1), int(st[3] >> 8)
            if info["root_moves"] == 0 or msc >= 50:
                winner = "draw" if msc >= 50 else seat[turn ^ 1]
                break
            if seat[turn] == "cand":
                s = trees[turn].search_ex(eng, seconds=seconds, key=1000 * g + plies, **cfg)
                po.append(s["playouts"])
                leaves.append(s["leaves"])
            else:
                trees[turn].search_ex(eng, iterations=opp_trials // 50, initial_batch=50, scale=0.0, max_batch=50, reps=1, depth=1,
                                      threads=1, key=5000 * g + plies)
            uct = seat[turn] == "cand" and cfg.get("policy") == 1
            m = trees[turn].robust_move(turn) if uct else trees[turn].best_move(turn)
            for t in trees:
                t.move(m)
            plies += 1
        score[winner] += 1
    p = (score["cand"] + 0.5 * score["draw"]) / games
    print(json.dumps({"candidate": name, "config": cfg, "seconds_per_move": seconds, "opponent": "host-like, %d trials/move in batches of 50" % opp_trials,
                      "games": games, "score": score, "score_rate": p, "wilson95": wilson(p, games),
                      "playouts_per_move": float(np.mean(po)), "leaf_selections_per_move": float(np.mean(leaves)),
                      "wall_s": round(time.time() - t_start, 1)}), flush=True)
