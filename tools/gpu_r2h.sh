#!/bin/bash
TAG=${1:-r02h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== match sanity"; timeout 300 shim/_ref/match_b200 b200 2 0.1 1 16384 16 2>&1 | tail -3 | cut -c1-400
timeout 200 shim/_ref/match_b200 hybrid 1 1 1 2>&1 | tail -2 | cut -c1-400
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 > $OUT/bench.json; python - <<PY
import json
j=json.load(open("$OUT/bench.json"))
print("value %.4e frac %.3f e2e %.4e packed %.4e heur %.4e (frac %.3f) dstart %.4e dlive %.4e search %.4e"%(j["value"],j["roofline"]["frac"],j["e2e"]["value"],j["e2e_packed"]["value"],j["heuristic"]["value"],j["heuristic"]["roofline"]["frac"],j["d_start"]["value"],j["d_live"]["value"],j["mcts_search"]["playouts_per_s_at_le_32_reps"]))
PY
