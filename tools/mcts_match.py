#!/usr/bin/env python3
"""BASELINE config 5 in a controlled form: B200 tree search vs an mcts_host-equivalent searcher.

Both players use the same tree policy (b2p_tree == the reference's GameTree, tests/test_tree.py) and the same
random-playout rules; they differ only in search budget:
  A  "b200"      : wall-clock budget per move on one B200 (large batches, `reps` playouts per selected leaf)
  B  "host-like" : a fixed number of trials per move in batches of 50, one playout per leaf = what the reference's
                   mcts_host (MCTSPlayer(50, 0, 7 s, HostPlayoutDriver), src/player.cpp:164-166) gets out of a
                   7 s move on the box's host cores (~7e4 trials) -- played on the GPU only
                   to save wall-clock; the search it performs is the same.
Colours alternate.  Prints one JSON line per game and a summary with a Wilson interval."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_ai_b200 as b  # noqa: E402
from oracle.pyoracle import START_PACKED  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=10)
ap.add_argument("--seconds", type=float, default=0.25)
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--reps", type=int, default=32)
ap.add_argument("--host-trials", type=int, default=70000)
args = ap.parse_args()
eng = b.Engine(devices=1)


def think(tree, who, key):
    if who == "b200":
        return tree.search(eng, seconds=args.seconds, initial_batch=args.batch, scale=0.02, reps=args.reps, key=key)
    return tree.search(eng, iterations=args.host_trials // 50, initial_batch=50, scale=0.0, reps=1, key=key)


score = {"b200": 0, "host-like": 0, "draw": 0}
playouts = {"b200": [], "host-like": []}
for g in range(args.games):
    seat = ("b200", "host-like") if g % 2 == 0 else ("host-like", "b200")
    trees = [b.Tree(START_PACKED), b.Tree(START_PACKED)]
    plies, t0 = 0, time.time()
    while True:
        info = trees[0].info()
        st = info["root_state"]
        turn, msc = int(st[3] & 1), int(st[3] >> 8)
        if info["root_moves"] == 0 or msc >= 50:
            winner = "draw" if msc >= 50 else seat[turn ^ 1]
            break
        playouts[seat[turn]].append(think(trees[turn], seat[turn], key=1000 * g + plies))
        m = trees[turn].best_move(turn)
        for t in trees:
            t.move(m)
        plies += 1
    score[winner] += 1
    print(json.dumps({"game": g, "p1": seat[0], "p2": seat[1], "winner": winner, "plies": plies, "seconds": round(time.time() - t0, 1)}), flush=True)

n = args.games
wins = score["b200"] + 0.5 * score["draw"]
p = wins / n
z = 1.96
lo = (p + z * z / (2 * n) - z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n))) / (1 + z * z / n)
hi = (p + z * z / (2 * n) + z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n))) / (1 + z * z / n)
print(json.dumps({"summary": score, "b200_score": p, "wilson95": [lo, hi], "b200_playouts_per_move": float(np.mean(playouts["b200"])),
                  "host_like_playouts_per_move": float(np.mean(playouts["host-like"])), "config": vars(args)}))
