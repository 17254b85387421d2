#!/bin/bash
# BASELINE config 5 on one box: 100 k ragged structures, Original / Tiny / Ultra-tiny, 1 / 2 / 4 / 8 GPUs
# (strong scaling of one host-side list), then one process driving all GPUs through device_ids.
# Usage (8-GPU box): tools/run_sweeps.sh [structures] ["1 2 4 8"] ["original tiny ultra_tiny"] [inprocess 0|1]
# JSON lines land in gpurun_out/sweep_c5_*.json
set -u
TOTAL=${1:-100000}
GPUS=${2:-"1 2 4 8"}
VARIANTS=${3:-"original tiny ultra_tiny"}
INPROCESS=${4:-1}
mkdir -p gpurun_out
python tools/make_sweep_cache.py $TOTAL /tmp/sweep_cache.npz
PORT=29600
for v in $VARIANTS; do
  for n in $GPUS; do
    PORT=$((PORT+1))
    if [ "$n" = "1" ]; then
      timeout 300 python bench.py --gpus 1 --mode sweep --variant $v --sweep-structures $TOTAL --sweep-cache /tmp/sweep_cache.npz > gpurun_out/sweep_c5_${v}_${n}gpu.json 2> gpurun_out/sweep_c5_${v}_${n}gpu.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $n --mode sweep --variant $v --sweep-structures $TOTAL --sweep-cache /tmp/sweep_cache.npz > gpurun_out/sweep_c5_${v}_${n}gpu.json 2> gpurun_out/sweep_c5_${v}_${n}gpu.err
    fi
    python -c "import json,sys; b=json.load(open('gpurun_out/sweep_c5_${v}_${n}gpu.json')); print('$v', $n, 'GPUs', round(b['value']), 'structures/s', b['phases_s'])" || tail -3 gpurun_out/sweep_c5_${v}_${n}gpu.err
  done
done
if [ "$INPROCESS" = "1" ]; then
  timeout 300 python tools/sweep_inprocess.py /tmp/sweep_cache.npz $TOTAL > gpurun_out/sweep_c5_inprocess.json 2> gpurun_out/sweep_c5_inprocess.err; cat gpurun_out/sweep_c5_inprocess.json
fi
