#!/bin/bash
# usage: tools/inflate_sweep.sh <workload> <entries> "<G W>" ...
wl=$1; n=$2; shift 2
for cfg in "$@"; do set -- $cfg
  OTZ_INFLATE_TILE=$1 OTZ_INFLATE_RING=$2 python bench.py --workload $wl --entries $n --steps 5 --no-cpu-baseline --e2e-steps 1 2>&1 | python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
try:
    d=json.loads(t[-1]); print('G=$1 W=$2 value %.1f GB/s decode %.3f ms'%(d['value'], d['roofline']['phase_ms']['decode']))
except Exception as e: print('G=$1 W=$2 FAILED', t[-3:])
"
done
