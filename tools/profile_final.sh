#!/bin/bash
# Final ncu evidence of a round (run under gpurun).  Every profiled command first exits 0 without ncu.
# usage: tools/profile_final.sh <round-tag>
set -x
T=${1:-r1}
mkdir -p gpurun_out
C2="python bench.py --workload c2 --entries 2500 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C1="python bench.py --workload c1 --entries 10000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C4="python bench.py --workload c4 --entries 2500 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C5="python bench.py --workload c5 --entries 1024 --steps 2 --warmup 3 --e2e-steps 1"
for W in C2 C1 C4 C5; do
  CMD=${!W}
  $CMD > gpurun_out/${T}_plain_$W.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_$W.csv $CMD > gpurun_out/${T}_ncu_$W.log 2>&1
done
$C2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_crc_chunks -s 3 -c 1 -o gpurun_out/${T}_prof_crc $C2 > gpurun_out/${T}_ncu_full_crc.log 2>&1
$C1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_inflate -s 3 -c 1 -o gpurun_out/${T}_prof_inflate $C1 > gpurun_out/${T}_ncu_full_inflate.log 2>&1
$C4 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_zstdref -s 3 -c 1 -o gpurun_out/${T}_prof_zstdref $C4 > gpurun_out/${T}_ncu_full_zstdref.log 2>&1
$C5 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_deflate_chunks -s 3 -c 1 -o gpurun_out/${T}_prof_deflate $C5 > gpurun_out/${T}_ncu_full_deflate.log 2>&1
ls -la gpurun_out | tail -20
