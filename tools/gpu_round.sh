#!/bin/bash
# One gpurun call: parity tests, smoke, microbenchmarks, bench, ncu launch list + full capture.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
(nproc; lscpu | head -20; free -g) > $OUT/host.txt 2>&1

echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== microbench"; timeout 300 python tools/microbench.py 2>&1 | tee $OUT/microbench.json
echo "== bench"; timeout 900 python bench.py --steps 8 --warmup 3 2>&1 | tail -3 | tee $OUT/bench.json
echo "== bench canonical order"; timeout 600 python bench.py --steps 4 --warmup 3 --order canonical --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_canonical.json
echo "== bench heuristic"; timeout 600 python bench.py --steps 4 --warmup 3 --mode heuristic --reps 8 --no-e2e 2>&1 | tail -1 | tee $OUT/bench_heuristic.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json

echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --reps 4 --no-cpu-baseline --no-e2e > $OUT/launches_run.log 2>&1
echo "== ncu full capture (playout kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:playout_lanes -s 4 -c 1 -f -o $OUT/prof_lanes \
  python bench.py --steps 1 --warmup 3 --reps 4 --no-cpu-baseline --no-e2e > $OUT/prof_run.log 2>&1
ls -la $OUT
