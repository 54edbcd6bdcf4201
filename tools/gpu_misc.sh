#!/bin/bash
# Sanitizer pass, the tests.sh-protocol sweep, small extras.  Usage: bash tools/gpu_misc.sh <tag>
TAG=${1:-misc}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "threads or errors" 2>&1 | tail -3
echo "== compute-sanitizer memcheck (smoke)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer_memcheck.txt 2>&1; echo "exit $?"; tail -4 $OUT/sanitizer_memcheck.txt
echo "== compute-sanitizer racecheck (heuristic + warp kernels, shared noise table)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python - > $OUT/sanitizer_racecheck.txt 2>&1 <<PY
import gpu_ai_b200 as b
e = b.Engine(devices=1)
st = e.gen_leaves(4096, key=2016)
for sched in (b.SCHED_THREAD, b.SCHED_WARP):
    e.run_packed(st, mode=b.MODE_HEURISTIC, sched=sched)
    e.run_packed(st, mode=b.MODE_RANDOM, sched=sched, order=b.ORDER_FAST)
print("racecheck workload done")
PY
echo "exit $?"; tail -4 $OUT/sanitizer_racecheck.txt
echo "== sweep (tests.sh protocol)"; timeout 900 python tools/sweep.py 1 > $OUT/sweep_1gpu.jsonl 2>&1; tail -3 $OUT/sweep_1gpu.jsonl
