#!/bin/bash
# A/B of the spline message kernels' launch shapes on C2 (one GPU): stage times per variant.
mkdir -p gpurun_out
for f in 0 1 2 3; do
  MLFFD_SPLINE_FWD=$f timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('fwd shape $f', round(b['value']), {k: round(v['ms_per_step'],3) for k,v in b['stages'].items() if k.startswith('message')})"
done
for r in 1 3 4 5; do
  MLFFD_SPLINE_BWD=$r timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('bwd shape $r', round(b['value']), {k: round(v['ms_per_step'],3) for k,v in b['stages'].items() if k.startswith('message')})"
done
