#!/bin/bash
# Final ncu evidence of a round (run under gpurun).  Every profiled command first exits 0 without ncu.
# usage: tools/profile_final.sh <round-tag>
# Leaves under gpurun_out/: <tag>_bench_*.json (bench lines), <tag>_launches_*.csv (per-launch durations),
# <tag>_ncu_*.json (summaries of one --set full capture per kernel, made here with tools/ncu_summary.py; the
# .ncu-rep files themselves are deleted — gpurun brings back at most 64 MiB).  Copy what should be judged to profiles/.
set -x
T=${1:-r1}
mkdir -p gpurun_out
C2="python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C2X="python bench.py --workload c2x --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C1="python bench.py --workload c1 --entries 10000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C4="python bench.py --workload c4 --entries 10000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C4Z="python bench.py --workload c4z --entries 2500 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C3W="python bench.py --workload c3w --entries 2000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C3="python bench.py --workload c3 --entries 2000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
C5="python bench.py --workload c5 --entries 1024 --steps 2 --warmup 3 --e2e-steps 1"
for W in C2 C2X C1 C4 C4Z C3W C3 C5; do
  CMD=${!W}
  $CMD > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_$W.csv $CMD > /dev/null 2>&1
done
full() {  # name, kernel regex, command, launches to skip
  $3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-3} -c 1 -o /tmp/${T}_prof_$1 $3 > gpurun_out/${T}_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py /tmp/${T}_prof_$1.ncu-rep > gpurun_out/${T}_ncu_$1.json 2>/dev/null
  python tools/ncu_lines.py /tmp/${T}_prof_$1.ncu-rep 30 > gpurun_out/${T}_lines_$1.txt 2>/dev/null
  rm -f /tmp/${T}_prof_$1.ncu-rep
}
full crc k_crc_chunks "$C2"
full storecopy k_store_copy "$C2X"
full tok k_inflate_tok "$C1"
full lz k_inflate_lz "$C1"
full zstdref k_zstdref "$C4"
full zseq k_zstd_seq "$C4Z"
full zlit k_zstd_lit "$C4Z"
full search k_block_search "$C3" 1
full lzseg k_inflate_lz "$C3" 4      # launches per pass: serial chain walk (empty), symbols (PAR), regular path -> the PAR one of pass 2
full segtok k_inflate_tok "$C3" 2     # per pass: segments, regular path -> the segment tokenizer of pass 2
full segtr k_seg_translate "$C3" 1
full deflate k_deflate_chunks "$C5"
ls -la gpurun_out | tail -40
