#!/bin/bash
TAG=${1:-r02n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 > $OUT/bench.json; python - <<PY
import json
j=json.load(open("$OUT/bench.json"))
print("value %.4e frac %.3f canonical %.4e e2e %.4e packed %.4e heur %.4e (frac %.3f) dstart %.4e dlive %.4e search %.4e"%(j["value"],j["roofline"]["frac"],j["canonical_order"]["playouts_per_s_per_gpu"],j["e2e"]["value"],j["e2e_packed"]["value"],j["heuristic"]["value"],j["heuristic"]["roofline"]["frac"],j["d_start"]["value"],j["d_live"]["value"],j["mcts_search"]["playouts_per_s_at_le_32_reps"]))
for r in j["mcts_search"]["configs"]: print("  search b%d reps %d depth %d: %.3e playouts/s, %.2e leaves/s, gpu_busy %.2f"%(r["initial_batch"],r["reps_per_leaf"],r["depth"],r["playouts_per_s"],r["leaf_selections_per_s"],r["gpu_busy"]))
PY
echo "== tree host bench"; timeout 300 python tools/tree_bench.py 1 8 16 2>&1 | grep -v 4096 > $OUT/tree_host_bench.jsonl; python - <<PY
import json
for l in open("$OUT/tree_host_bench.jsonl"):
    j=json.loads(l); print(j["interface"][:30], j.get("threads"), j.get("exact"), j["batch"], "%.2e"%j["leaves_per_s"])
PY
echo "== sweep"; timeout 600 python tools/sweep.py 1 > $OUT/sweep_1gpu.jsonl 2>&1; grep device_single $OUT/sweep_1gpu.jsonl | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print(j['n'], '%.3e'%j['playouts_per_s'], '%.3f ms'%(1e3*j['best_s']))"
echo "== selfplay (UCT candidates)"; timeout 900 python tools/selfplay.py 8 0.1 100000 > $OUT/selfplay_uct.jsonl 2>&1; cut -c1-330 $OUT/selfplay_uct.jsonl
