#!/bin/bash
# BASELINE configs 4/5: the UNMODIFIED reference MCTS players (src/mcts.cpp, src/player.cpp, 7 s per move)
# on top of the B200 drop-in, against the reference's own mcts_host.  One game per pairing and colour.
TAG=${1:-games}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for pair in "mcts_device_coarse mcts_host" "mcts_host mcts_device_coarse" ; do
  set -- $pair
  echo "== game_test $1 (P1) vs $2 (P2)" | tee -a $OUT/games.txt
  ( time timeout 1500 shim/_ref/run_ai_b200 -m game_test -n 1 -1 $1 -2 $2 ) 2>&1 | tail -12 | tee -a $OUT/games.txt
done
