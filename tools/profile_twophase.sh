#!/bin/bash
# usage: tools/profile_twophase.sh <tag> [entries]   (run under gpurun) — ncu --set full of k_inflate_tok and k_inflate_lz on C1
set -x
T=$1; N=${2:-10000}
mkdir -p gpurun_out
CMD="python bench.py --workload c1 --entries $N --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/${T}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_inflate_tok -s 3 -c 1 -o gpurun_out/${T}_tok $CMD > gpurun_out/${T}_ncu_tok.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_inflate_lz -s 3 -c 1 -o gpurun_out/${T}_lz $CMD > gpurun_out/${T}_ncu_lz.log 2>&1
tail -1 gpurun_out/${T}_plain.log | cut -c1-200
