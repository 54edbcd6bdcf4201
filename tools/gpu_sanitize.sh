#!/bin/bash
# compute-sanitizer passes over every kernel variant (memcheck, racecheck, synccheck).  Usage: bash tools/gpu_sanitize.sh <tag>
TAG=${1:-sanitize}; OUT=gpurun_out/$TAG; mkdir -p $OUT
cat > /tmp/san_workload.py <<PY
import numpy as np, gpu_ai_b200 as b
e = b.Engine(devices=1)
st = e.gen_leaves(6000, key=2016)
for sched in (b.SCHED_THREAD, b.SCHED_WARP):
    for mode, order in ((b.MODE_RANDOM, b.ORDER_FAST), (b.MODE_RANDOM, b.ORDER_CANONICAL), (b.MODE_HEURISTIC, b.ORDER_CANONICAL)):
        e.run_packed(st, reps=2, mode=mode, sched=sched, order=order, want_plies=True, want_final=True)
        e.run_packed(st, reps=2, mode=mode, sched=sched, order=order)
        e.run_counts(st, reps=3, mode=mode, sched=sched, order=order)
    e.run_packed(st, max_plies=7, sched=sched, want_final=True)
e.genmoves(st, 48)
t = b.Tree(st[5]); t.search(e, iterations=5, initial_batch=300, reps=4)
t = b.Tree(st[7]); t.search_ex(e, iterations=6, initial_batch=3000, reps=4, depth=2, policy=1)
from gpu_ai_b200 import engine as E
e3 = b.Engine(devices=[0, 0, 0])
s776 = E.unpack776(st); e3.run_states776(np.concatenate([s776] * 6)); e.run_states776(s776[:700], mode=b.MODE_HEURISTIC, sched=b.SCHED_AUTO)
print("sanitizer workload done")
PY
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  echo "== $tool"; PYTHONPATH=$PWD timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_workload.py > $OUT/sanitizer_$tool.txt 2>&1; echo "exit $?"; tail -3 $OUT/sanitizer_$tool.txt
done
