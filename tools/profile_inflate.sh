#!/bin/bash
# usage: tools/profile_inflate.sh <tag> <entries> <G> <W>
set -x
mkdir -p gpurun_out
export OTZ_INFLATE_TILE=$3 OTZ_INFLATE_RING=$4
CMD="python bench.py --workload c1 --entries $2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plain_$1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_inflate -s 3 -c 1 -o gpurun_out/prof_inflate_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log
