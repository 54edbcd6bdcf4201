#!/bin/bash
# Quick GPU iteration: parity tests, bench (fast + canonical), one ncu full capture of the playout kernel.
TAG=${1:-iter}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench.json
echo "== bench canonical"; timeout 600 python bench.py --steps 4 --warmup 3 --order canonical --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_canonical.json
echo "== bench heuristic"; timeout 600 python bench.py --steps 4 --warmup 3 --mode heuristic --reps 8 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_heuristic.json
echo "== ncu full capture (playout kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:playout_lanes -s 4 -c 1 -f -o $OUT/prof_lanes \
  python bench.py --steps 1 --warmup 3 --reps 16 --no-cpu-baseline --no-e2e > $OUT/prof_run.log 2>&1
tail -2 $OUT/prof_run.log | cut -c1-300
if [ "$2" == "extras" ]; then
  echo "== drop-in run_ai (shim)"; 
  (timeout 300 shim/_ref/run_ai_b200 -m playout_test -n 200000 -1 device_single -2 host; 
   timeout 300 shim/_ref/run_ai_b200 -m playout_test -n 200000 -1 device_heuristic -2 host_heuristic;
   timeout 300 shim/_ref/run_ai_b200 -m playout_test -n 200000 -1 device_multiple -2 device_coarse;
   timeout 300 shim/_ref/run_ai_b200 -m gen_moves_test -n 20000) 2>&1 | tee $OUT/run_ai_b200.txt | tail -40
  echo "== reference kernels recompiled for sm_100a (secondary comparison)";
  (timeout 120 oracle/_ref/run_ai_ref -m playout_test -n 100000 -1 device_single -2 device_coarse;
   timeout 120 oracle/_ref/run_ai_ref -m playout_test -n 100000 -1 device_multiple -2 device_heuristic) 2>&1 | tee $OUT/run_ai_ref.txt | tail -40
fi
ls -la $OUT
if [ "$2" == "sched" ] || [ "$3" == "sched" ]; then
  echo "== thread vs warp scheduling"; timeout 600 python tools/sched_compare.py 2>&1 | tee $OUT/sched_compare.jsonl
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:playout_warp -s 2 -c 1 -f -o $OUT/prof_warp \
    python - > $OUT/prof_warp_run.log 2>&1 <<PY
import torch, gpu_ai_b200 as b
eng=b.Engine(devices=[0]); dev=torch.device("cuda",0); N=1<<17
s=torch.empty((N,4),dtype=torch.int32,device=dev); eng.gen_leaves_device(N,s.data_ptr(),key=2016)
w=torch.empty(N,dtype=torch.int8,device=dev); c=torch.zeros(4,dtype=torch.int64,device=dev)
for i in range(4):
    eng.run_packed_device(s.data_ptr(),N,reps=1,key=i,mode=b.MODE_RANDOM,sched=b.SCHED_WARP,order=b.ORDER_FAST,d_winners=w.data_ptr(),d_counters=c.data_ptr(),stream=0)
torch.cuda.synchronize()
PY
fi
