#!/bin/bash
# round 2, run B: tests + the full default bench line + reference arm + search throughput table
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== bench (default command)"; ( time timeout 900 python bench.py ) 2>&1 | tail -5 | tee $OUT/bench_full.txt | cut -c1-300; grep '^{' $OUT/bench_full.txt > $OUT/bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-600
echo "== tree host bench"; timeout 300 python tools/tree_bench.py 1 8 16 > $OUT/tree_host_bench.jsonl 2>&1; cut -c1-220 $OUT/tree_host_bench.jsonl
echo "== sweep"; timeout 600 python tools/sweep.py 1 > $OUT/sweep_1gpu.jsonl 2>&1; grep device_single $OUT/sweep_1gpu.jsonl | cut -c1-160
