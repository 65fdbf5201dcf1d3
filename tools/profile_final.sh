#!/bin/bash
# Final ncu evidence of a round (run under gpurun).  Every profiled command first exits 0 without ncu.
# usage: tools/profile_final.sh <round-tag>
# Leaves under gpurun_out/: <tag>_plain_*.log (bench lines), <tag>_launches_*.csv (per-launch durations),
# <tag>_prof_*.ncu-rep (--set full captures); copy the summaries you want judged into profiles/.
set -x
T=${1:-r1}
mkdir -p gpurun_out
C2="python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C1="python bench.py --workload c1 --entries 10000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C4="python bench.py --workload c4 --entries 10000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C4Z="python bench.py --workload c4z --entries 2500 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C3W="python bench.py --workload c3w --entries 2000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C5="python bench.py --workload c5 --entries 1024 --steps 2 --warmup 3 --e2e-steps 1"
for W in C2 C1 C4 C4Z C3W C5; do
  CMD=${!W}
  $CMD > gpurun_out/${T}_plain_$W.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_$W.csv $CMD > gpurun_out/${T}_ncu_$W.log 2>&1
done
full() {  # name, kernel regex, command
  $3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/${T}_prof_$1 $3 > gpurun_out/${T}_ncu_full_$1.log 2>&1
}
full crc k_crc_chunks "$C2"
full tok k_inflate_tok "$C1"
full lz k_inflate_lz "$C1"
full zstdref k_zstdref "$C4"
full zstd "k_zstd$" "$C4Z"
full deflate k_deflate_chunks "$C5"
ls -la gpurun_out | tail -30
