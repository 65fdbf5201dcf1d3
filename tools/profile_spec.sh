#!/bin/bash
# --set full captures of the speculative tokenizer (run under gpurun).  usage: tools/profile_spec.sh <tag>
T=${1:-r2}
mkdir -p gpurun_out
full() {  # name, kernel regex, command, launches to skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-2} -c 1 -o /tmp/${T}_prof_$1 $3 > gpurun_out/${T}_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py /tmp/${T}_prof_$1.ncu-rep > gpurun_out/${T}_ncu_$1.json 2>/dev/null
  python tools/ncu_lines.py /tmp/${T}_prof_$1.ncu-rep 45 > gpurun_out/${T}_lines_$1.txt 2>/dev/null
  rm -f /tmp/${T}_prof_$1.ncu-rep
}
C1="python bench.py --workload c1 --entries 10000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
C3="python bench.py --workload c3 --entries 2000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
for K in ${KERNELS:-spec1 spec4 lz}; do
  case $K in
    spec1) full spec1 "k_inflate_spec" "$C1" 2 ;;
    spec4) full spec4 "k_inflate_spec" "$C3" 2 ;;   # per pass: spec<4>, spec<1>
    lz) full lz "k_inflate_lz" "$C1" 2 ;;
    par) full par "k_inflate_lz" "$C3" 4 ;;     # per pass: serial chain walk (empty), symbols, regular -> the symbol executor of pass 2
    lz3) full lz3 "k_inflate_lz" "$C3" 5 ;;     # the byte executor of pass 2 on configs[2]
    spec4w) full spec4w "k_inflate_spec" "python bench.py --workload c3 --entries 1250 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0" 2 ;;
    win) full win "k_seg_window" "$C3" 1 ;;
    tr) full tr "k_seg_translate" "$C3" 1 ;;
  esac
done
