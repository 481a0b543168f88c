#!/bin/bash
# tools/run_8gpu.sh - the multi-GPU measurements of round 2 in one gpurun --gpus 8 call (results under gpurun_out/, copied to profiles/):
#   1. BASELINE configs[4] (5000 views x 5000 segments, ring +-5) through L3DPP::Line3D + setShard on 8 GPUs, diffusion on
#   2. BASELINE configs[3] (1000 x 3000) STRONG scaling 1/2/4/8: whole pipeline (run_dist_pipeline.py) and matching alone (bench.py --scaling strong)
#   3. the 2-process NCCL parity test
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
p=29600
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/r02_8gpu_devices.csv 2>&1
# 1. cfg5 on 8 GPUs (+ bit check sharded == unsharded on a 48-view sub-ring)
timeout 900 $TR --nproc-per-node 8 --master-port $((p++)) tools/run_dist_pipeline.py --views 5000 --segs 5000 --ring 5 --diffusion 1 --check-views 48 --reps 2 \
    > $O/r02_cfg5_8gpu.json 2> $O/r02_cfg5_8gpu.err
# 2a. strong scaling of the whole pipeline
for n in 1 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port $((p++)) tools/run_dist_pipeline.py --views 1000 --segs 3000 --ring 5 --diffusion 1 --reps 3 \
      > $O/r02_strong_pipeline_$n.json 2> $O/r02_strong_pipeline_$n.err
done
# 2b. strong scaling of the matching step (bench contract, one JSON line each; lean: no roofline legs / baselines)
export L3D_BENCH_LEAN=1
for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 900 python bench.py --gpus 1 --scaling strong --steps 5 --warmup 3 > $O/r02_strong_$n.json 2> $O/r02_strong_$n.err
  else timeout 900 $TR --nproc-per-node $n --master-port $((p++)) bench.py --gpus $n --scaling strong --steps 5 --warmup 3 > $O/r02_strong_$n.json 2> $O/r02_strong_$n.err; fi
done
unset L3D_BENCH_LEAN
# 3. two processes over NCCL: sharded == single GPU, bit for bit
timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu > $O/r02_pytest_dist_8gpubox.log 2>&1
tail -c 400 $O/r02_cfg5_8gpu.json; tail -n 3 $O/r02_cfg5_8gpu.err
for n in 1 2 4 8; do python - <<PY
import json
try:
    d=json.loads(open("$O/r02_strong_pipeline_$n.json").read().strip().splitlines()[-1]); print("pipeline", $n, d["matchImages_s_max"], d["reconstruct_s_max"], d["ms_score_max"], d["ms_exchange_max"], d["identical_result_on_every_rank"])
    d=json.loads(open("$O/r02_strong_$n.json").read().strip().splitlines()[-1]); print("match", $n, d["value"], d["ms_per_step"], d["e2e"]["value"])
except Exception as e: print("n=$n", e)
PY
done
tail -n 2 $O/r02_pytest_dist_8gpubox.log
