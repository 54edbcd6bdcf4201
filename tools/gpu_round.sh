#!/bin/bash
# End-of-round evidence run (one gpurun call, 1 GPU): parity tests, smoke, microbenchmarks, bench lines,
# reference arm, ncu launch list + full capture of the dominant kernel AT THE BENCH CONFIGURATION.
# Usage (repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
(nproc; lscpu | head -20; free -g) > $OUT/host.txt 2>&1

echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== microbench"; timeout 300 python tools/microbench.py > $OUT/microbench.json 2>&1; tail -3 $OUT/microbench.json
echo "== bench (default command)"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json | cut -c1-400
echo "== bench canonical order"; timeout 600 python bench.py --steps 6 --warmup 3 --order canonical --no-cpu-baseline --no-e2e 2>&1 | tail -1 > $OUT/bench_canonical.json
echo "== bench heuristic"; timeout 600 python bench.py --steps 6 --warmup 3 --mode heuristic --reps 8 2>&1 | tail -1 > $OUT/bench_heuristic.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-300

echo "== ncu launch list (same command as the bench, fewer steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/launches_run.log 2>&1
echo "== ncu full capture: dominant kernel at the bench configuration (2^20 leaves x 32 reps, fast order)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:playout_lanes -s 4 -c 1 -f -o $OUT/prof_lanes_bench_config \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/prof_run.log 2>&1
echo "== ncu full capture: heuristic kernel"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:playout_lanes -s 4 -c 1 -f -o $OUT/prof_lanes_heuristic \
  python bench.py --steps 1 --warmup 3 --mode heuristic --reps 8 --no-cpu-baseline --no-e2e > $OUT/prof_heur_run.log 2>&1
ls -la $OUT
echo "== scheduler comparison"; timeout 600 python tools/sched_compare.py > $OUT/sched_compare.jsonl 2>&1; tail -2 $OUT/sched_compare.jsonl
echo "== tests.sh-protocol sweep"; timeout 600 python tools/sweep.py 1 > $OUT/sweep_1gpu.jsonl 2>&1; tail -1 $OUT/sweep_1gpu.jsonl
echo "== MCTS search throughput"; timeout 300 python tools/mcts_bench.py > $OUT/mcts_search_throughput.jsonl 2>&1; tail -2 $OUT/mcts_search_throughput.jsonl
echo "== MCTS match (8 games)"; timeout 900 python tools/mcts_match.py --games 8 > $OUT/mcts_match.jsonl 2>&1; tail -1 $OUT/mcts_match.jsonl
echo "== drop-in run_ai"; (timeout 300 shim/_ref/run_ai_b200 -m playout_test -n 200000 -1 device_single -2 host; timeout 300 shim/_ref/run_ai_b200 -m gen_moves_test -n 20000) > $OUT/run_ai_b200.txt 2>&1; tail -3 $OUT/run_ai_b200.txt
ls -la $OUT
