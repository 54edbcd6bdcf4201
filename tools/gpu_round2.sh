#!/bin/bash
# End-of-round-2 evidence run (one gpurun call, 1 GPU): parity tests, smoke, microbenchmarks, bench lines, reference arm,
# ncu launch list + full captures of the dominant kernels AT THE BENCH CONFIGURATION, sanitizer, sweeps.
TAG=${1:-r02_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
(nproc; lscpu | head -20; free -g) > $OUT/host.txt 2>&1

echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== microbench"; timeout 300 python tools/microbench.py > $OUT/microbench.json 2>&1; tail -3 $OUT/microbench.json
echo "== bench (default command)"; ( time timeout 900 python bench.py ) 2>&1 | tail -5 > $OUT/bench_full.txt; grep '^{' $OUT/bench_full.txt > $OUT/bench.json; cut -c1-300 $OUT/bench.json; grep real $OUT/bench_full.txt
echo "== bench canonical order"; timeout 600 python bench.py --steps 6 --warmup 3 --order canonical --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 > $OUT/bench_canonical.json
echo "== bench heuristic (own line)"; timeout 600 python bench.py --steps 6 --warmup 3 --mode heuristic --reps 16 --no-extras 2>&1 | tail -1 > $OUT/bench_heuristic.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-300

echo "== ncu launch list (same command as the bench, fewer steps, no host-side sections)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/launches_run.log 2>&1
echo "== ncu full capture: dominant kernel at the bench configuration (2^20 leaves x 64 reps, fast order)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:playout_lanes -s 4 -c 1 -f -o $OUT/prof_lanes_bench_config \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_run.log 2>&1
echo "== ncu full capture: heuristic kernel"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:playout_lanes -s 4 -c 1 -f -o $OUT/prof_lanes_heuristic \
  python bench.py --steps 1 --warmup 3 --mode heuristic --reps 16 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_heur_run.log 2>&1
ls -la $OUT
echo "== tests.sh-protocol sweep"; timeout 600 python tools/sweep.py 1 > $OUT/sweep_1gpu.jsonl 2>&1; tail -1 $OUT/sweep_1gpu.jsonl | cut -c1-200
echo "== MCTS search throughput"; timeout 300 python tools/mcts_bench.py 1 1.0 > $OUT/mcts_search_throughput.jsonl 2>&1; tail -2 $OUT/mcts_search_throughput.jsonl | cut -c1-300
echo "== drop-in run_ai"; (B2P_ROUTING_REPORT=1 timeout 300 shim/_ref/run_ai_b200 -m playout_test -n 200000 -1 device_single -2 host; B2P_ROUTING_REPORT=1 timeout 300 shim/_ref/run_ai_b200 -m playout_test -n 200000 -1 hybrid -2 optimal; timeout 300 shim/_ref/run_ai_b200 -m gen_moves_test -n 20000) > $OUT/run_ai_b200.txt 2>&1; tail -4 $OUT/run_ai_b200.txt
echo "== sanitizer (memcheck)"; bash tools/gpu_sanitize.sh $TAG 2>&1 | tail -12
ls -la $OUT
