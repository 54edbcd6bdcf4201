#!/bin/bash
# multi-GPU validation (run with gpurun --gpus N): in-process multi-device tests on distinct GPUs, torchrun bench at N, reference arm under torchrun
N=${1:-2}; TAG=${2:-r02k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L | tee $OUT/gpus.txt
echo "== pytest (multi-device + boundary tests)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device or every_batch_size or tree_search or run_counts" 2>&1 | tail -4 | tee $OUT/pytest_multi.txt
echo "== torchrun bench N=$N"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 ) 2>&1 | grep -E '^\{|^real' > $OUT/bench_n$N.txt; grep '^{' $OUT/bench_n$N.txt > $OUT/bench_n$N.json
python - <<PY
import json
j=json.load(open("$OUT/bench_n$N.json"))
print("N=%d value %.4e e2e %.4e packed %.4e shard_inv %s inproc %s"%(j["n_gpus"],j["value"],j["e2e"]["value"],j["e2e_packed"]["value"],j.get("shard_invariance"),j.get("in_process_multi_device")))
for r in j["mcts_search"]["configs"]: print("  search reps %d depth %d: %.3e playouts/s, %.2e leaves/s, gpu_busy %.2f"%(r["reps_per_leaf"],r["depth"],r["playouts_per_s"],r["leaf_selections_per_s"],r["gpu_busy"]))
PY
grep real $OUT/bench_n$N.txt
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>&1 | grep '^{' | tee $OUT/bench_reference_n$N.json | cut -c1-700
echo "== sweep with $N devices"; timeout 600 python tools/sweep.py $N > $OUT/sweep_${N}gpu.jsonl 2>&1; grep device_single $OUT/sweep_${N}gpu.jsonl | cut -c1-150
echo "== search over $N devices"; timeout 300 python tools/mcts_bench.py $N 1.0 > $OUT/mcts_${N}gpu.jsonl 2>&1; python - <<PY
import json
for l in open("$OUT/mcts_${N}gpu.jsonl"):
    try: j=json.loads(l)
    except Exception: print(l); continue
    print("b%-8d reps%-4d d%d  %.3e po/s  %.2e leaves/s  busy %.2f sel %.2f wait %.2f"%(j['initial_batch'],j['reps'],j['depth'],j['playouts_per_s'],j['leaf_selections_per_s'],j['gpu_busy'],j['select_s'],j['wait_s']))
PY
