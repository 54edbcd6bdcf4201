#!/bin/bash
# BASELINE config 5: literal games against the reference's own mcts_host player (shim/match_main.cpp)
TAG=${1:-r02m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== b200 tree player 1 s/move vs mcts_host 1 s/move (equal budgets)"
timeout 1500 shim/_ref/match_b200 b200 14 1.0 1 2048 16 > $OUT/match_b200_1s_vs_mcts_host_1s.jsonl 2>&1; tail -1 $OUT/match_b200_1s_vs_mcts_host_1s.jsonl | cut -c1-400
echo "== b200 tree player 0.1 s/move vs mcts_host 1 s/move (a tenth of the time)"
timeout 1500 shim/_ref/match_b200 b200 20 0.1 1 2048 16 > $OUT/match_b200_0.1s_vs_mcts_host_1s.jsonl 2>&1; tail -1 $OUT/match_b200_0.1s_vs_mcts_host_1s.jsonl | cut -c1-400
echo "== reference mcts_hybrid preset on the drop-in, 1 s/move vs mcts_host 1 s/move"
B2P_ROUTING_REPORT=1 timeout 900 shim/_ref/match_b200 hybrid 8 1 1 > $OUT/match_mcts_hybrid_dropin_vs_mcts_host.jsonl 2>&1; tail -3 $OUT/match_mcts_hybrid_dropin_vs_mcts_host.jsonl | cut -c1-400
