#!/bin/bash
# A/B of library variants: heuristic + random + canonical bench values (B2P_LIB_PATH), quick parity per variant
TAG=${1:-ab}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for v in base "$@"; do
  if [ "$v" == "base" ]; then unset B2P_LIB_PATH; else export B2P_LIB_PATH=$PWD/gpu_ai_b200/libb2p_$v.so; fi
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or oracle_large" 2>&1 | tail -1
  timeout 600 python bench.py --steps 6 --warmup 3 --mode heuristic --reps 8 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 > $OUT/bench_${v}_heur.json
  python -c "import json;d=json.load(open('$OUT/bench_${v}_heur.json'));print('$v heuristic %.4e playouts/s frac %.3f'%(d['value'],d['roofline']['frac']))"
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 > $OUT/bench_${v}_rand.json
  python -c "import json;d=json.load(open('$OUT/bench_${v}_rand.json'));print('$v random %.4e playouts/s frac %.3f canonical %.4e'%(d['value'],d['roofline']['frac'],d['canonical_order']['playouts_per_s_per_gpu']))"
done
