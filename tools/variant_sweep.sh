#!/bin/bash
# A/B of library builds (build_variants/libmlffd_*.so, selected through MLFFD_LIB) and launch shapes on C2,
# one GPU: value and message-stage times per variant.  Usage: tools/variant_sweep.sh "s0p0 s1p1" "0 3" "0 1"
mkdir -p gpurun_out
VARIANTS=${1:-"s0p0 s1p0 s0p1 s1p1 s2p1"}
FWDS=${2:-"0"}
BWDS=${3:-"0"}
for v in $VARIANTS; do for f in $FWDS; do for r in $BWDS; do
  MLFFD_LIB=$PWD/build_variants/libmlffd_$v.so MLFFD_SPLINE_FWD=$f MLFFD_SPLINE_BWD=$r \
    timeout 150 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/variant_err.log | python -c "
import json,sys
try:
    b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v fwd$f bwd$r', round(b['value']), {k: round(x['ms_per_step'],3) for k,x in b['stages'].items() if k.startswith('message')}, flush=True)
except Exception as ex:
    print('$v fwd$f bwd$r FAILED', ex, open('gpurun_out/variant_err.log').read()[-400:])"
done; done; done | tee -a gpurun_out/variant_sweep.log
