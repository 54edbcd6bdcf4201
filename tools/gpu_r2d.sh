#!/bin/bash
# round 2, run D: heuristic kernel v2 -- parity, A/B of register caps, ncu capture
TAG=${1:-r02d}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (heuristic + parity subset)"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
for v in base "$@"; do
  if [ "$v" == "base" ]; then unset B2P_LIB_PATH; else export B2P_LIB_PATH=$PWD/gpu_ai_b200/libb2p_$v.so; fi
  timeout 600 python bench.py --steps 6 --warmup 3 --mode heuristic --reps 8 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 > $OUT/bench_${v}_heur.json
  python -c "import json;d=json.load(open('$OUT/bench_${v}_heur.json'));print('$v heuristic %.4e playouts/s frac %.3f'%(d['value'],d['roofline']['frac']))"
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 > $OUT/bench_${v}_rand.json
  python -c "import json;d=json.load(open('$OUT/bench_${v}_rand.json'));print('$v random %.4e playouts/s frac %.3f canonical %.4e'%(d['value'],d['roofline']['frac'],d['canonical_order']['playouts_per_s_per_gpu']))"
done
unset B2P_LIB_PATH
echo "== ncu full capture: heuristic kernel"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:playout_lanes -s 4 -c 1 -f -o $OUT/prof_lanes_heuristic \
  python bench.py --steps 1 --warmup 3 --mode heuristic --reps 8 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_heur_run.log 2>&1
ls -la $OUT
