#!/bin/bash
# BASELINE config 5: literal games against the reference's own mcts_host player (shim/match_main.cpp).
# The B200 player searches with B2P_POLICY_UCT (fixed batches of 2048 leaves x 16 playouts, two batches in flight).
TAG=${1:-r02p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== b200 tree player 0.1 s/move vs mcts_host 1 s/move (a tenth of the time)"
timeout 1500 shim/_ref/match_b200 b200 20 0.1 1 2048 16 0 1 > $OUT/match_b200_uct_0.1s_vs_mcts_host_1s.jsonl 2>&1; tail -1 $OUT/match_b200_uct_0.1s_vs_mcts_host_1s.jsonl | cut -c1-400
echo "== b200 tree player 1 s/move vs mcts_host 1 s/move (equal budgets)"
timeout 1200 shim/_ref/match_b200 b200 10 1.0 1 2048 16 0 1 > $OUT/match_b200_uct_1s_vs_mcts_host_1s.jsonl 2>&1; tail -1 $OUT/match_b200_uct_1s_vs_mcts_host_1s.jsonl | cut -c1-400
