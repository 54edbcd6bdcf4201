#!/bin/bash
# ncu full capture of one playout kernel.  Usage: bash tools/gpu_prof.sh <tag> <mode: random|heuristic> <order> <reps>
TAG=${1:-prof}; MODE=${2:-random}; ORDER=${3:-fast}; REPS=${4:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:playout_lanes -s 4 -c 1 -f -o $OUT/prof_${MODE}_${ORDER} \
  python bench.py --steps 1 --warmup 3 --reps $REPS --mode $MODE --order $ORDER --no-cpu-baseline --no-e2e > $OUT/prof_${MODE}_${ORDER}.log 2>&1
tail -1 $OUT/prof_${MODE}_${ORDER}.log | cut -c1-200
