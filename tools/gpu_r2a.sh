#!/bin/bash
# round 2, run A: parity tests on the new boundary pipeline, tests.sh-protocol sweep, bench
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
(nproc; free -g | head -2) > $OUT/host.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== sweep"; timeout 600 python tools/sweep.py 1 > $OUT/sweep_1gpu.jsonl 2>&1; grep device_single $OUT/sweep_1gpu.jsonl | cut -c1-160
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench.json | cut -c1-600
