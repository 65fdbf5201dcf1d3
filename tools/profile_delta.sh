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
for W in ${WORKLOADS:-C1 C3 C3W C4Z C5}; do
  CMD=${!W}
  $CMD > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_$W.csv ${CMD/--steps 5/--steps 1} > /dev/null 2>&1
done
if [ "$2" = "full" ]; then
  # one --set full capture of the kernels of the segmented path (after the un-profiled run above exited 0)
  full() {  # name, kernel regex, command, launches to skip
    ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-3} -c 1 -o /tmp/${T}_prof_$1 $3 > gpurun_out/${T}_ncu_full_$1.log 2>&1
    python tools/ncu_summary.py /tmp/${T}_prof_$1.ncu-rep > gpurun_out/${T}_ncu_$1.json 2>/dev/null
    python tools/ncu_lines.py /tmp/${T}_prof_$1.ncu-rep 30 > gpurun_out/${T}_lines_$1.txt 2>/dev/null
    rm -f /tmp/${T}_prof_$1.ncu-rep
  }
  C3S="python bench.py --workload c3 --entries 2000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
  full search k_block_search "$C3S" 1
  full verify k_block_verify "$C3S" 1
  full lzseg k_inflate_lz "$C3S" 4      # launches per pass: serial chain walk (empty), symbols, regular path -> the symbol executor of pass 2
  full segtok k_inflate_tok "$C3S" 3    # per pass: regular path, segments -> the segment tokenizer of pass 2
fi
