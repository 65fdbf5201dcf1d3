#!/bin/bash
# Bench lines and per-launch durations of the workloads a late change touched (run under gpurun).
# usage: tools/profile_delta.sh <tag>
set -x
T=${1:-r1h}
mkdir -p gpurun_out
C1="python bench.py --workload c1 --entries 10000 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3"
C3="python bench.py --workload c3 --entries 2000 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3"
C3W="python bench.py --workload c3w --entries 2000 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3"
C4Z="python bench.py --workload c4z --entries 2500 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3"
C5="python bench.py --workload c5 --entries 1024 --steps 3 --warmup 3 --e2e-steps 1"
for W in C1 C3 C3W C4Z C5; do
  CMD=${!W}
  $CMD > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_$W.csv ${CMD/--steps 5/--steps 1} > /dev/null 2>&1
done
