#!/bin/bash
# A/B experiment builds: bench value per variant (B2P_LIB_PATH), parity check for each.
TAG=${1:-ab}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
python tools/microbench.py > $OUT/microbench.json 2>&1; grep -A1 -E '"(IMAD.HI|LOP3 \+ mul|LOP3 \+ SHF|LOP3)' $OUT/microbench.json | grep -E 'IMAD|LOP3|per_sm' | paste - - | cut -c1-160
for v in base "$@"; do
  if [ "$v" == "base" ]; then unset B2P_LIB_PATH; else export B2P_LIB_PATH=$PWD/gpu_ai_b200/libb2p_$v.so; fi
  echo "== variant $v"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or oracle_large" 2>&1 | tail -1
  for order in fast canonical; do
    timeout 600 python bench.py --steps 8 --warmup 3 --order $order --no-cpu-baseline --no-e2e 2>&1 | tail -1 > $OUT/bench_${v}_$order.json
    python -c "import json;d=json.load(open('$OUT/bench_${v}_$order.json'));print('$v $order %.4e playouts/s single_pass %.3f ms'%(d['value'],d['single_pass']['ms']))"
  done
  timeout 600 python bench.py --steps 4 --warmup 3 --mode heuristic --reps 8 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > $OUT/bench_${v}_heur.json
  python -c "import json;d=json.load(open('$OUT/bench_${v}_heur.json'));print('$v heuristic %.4e playouts/s'%(d['value']))"
done
