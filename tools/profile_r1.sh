#!/bin/bash
# ncu evidence for round 1 (run under gpurun; every profiled command first exits 0 without ncu)
set -x
mkdir -p gpurun_out
C1="python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C2="python bench.py --workload c2 --entries 2000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$C1 > gpurun_out/plain_c1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c1.csv $C1 > gpurun_out/ncu_c1.log 2>&1
$C1 > gpurun_out/plain_c1b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_inflate -s 3 -c 1 -o gpurun_out/prof_inflate $C1 > gpurun_out/ncu_c1_full.log 2>&1
$C2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv $C2 > gpurun_out/ncu_c2.log 2>&1
$C2 > gpurun_out/plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_crc_chunks -s 3 -c 1 -o gpurun_out/prof_crc $C2 > gpurun_out/ncu_c2_full.log 2>&1
ls -la gpurun_out
