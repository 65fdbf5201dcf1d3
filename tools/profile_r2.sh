#!/bin/bash
# Round-2 ncu evidence (run under gpurun).  Every profiled command first exits 0 without ncu.
#   usage: tools/profile_r2.sh <tag>
# Leaves under gpurun_out/: <tag>_bench_default.json (the default bench line, secondary workloads included),
# <tag>_launches_C3.csv (per-launch durations of one pass of the default workload, configs[2] as named),
# <tag>_ncu_<kernel>.json / <tag>_lines_<kernel>.txt (summary + per-line stalls of one --set full capture per kernel of
# the decode phase, made here with tools/ncu_summary.py / ncu_lines.py; the .ncu-rep files are deleted — gpurun brings
# back at most 64 MiB), <tag>_traffic_C3.json (DRAM bytes of every decode-phase kernel of one pass: bench.py's
# roofline.traffic), <tag>_sass_*.txt (SASS excerpts: bulk copy / mbarrier in k_seg_translate, the CRC fold loop).
T=${1:-r2}
mkdir -p gpurun_out
DEF="python bench.py --steps 5 --warmup 3"
C3="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1"
$DEF > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
OTZ_PIPE_TRACE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > /dev/null 2> gpurun_out/${T}_pipe_trace.txt
$C3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_C3.csv $C3 > /dev/null 2>&1
# one --set full capture of every kernel of the decode phase of the LAST warm-up pass (12 launches per pass: skip 3 passes)
ncu --set full --clock-control none --import-source on -k regex:'k_inflate|k_seg|k_crc_chunks' -s 27 -c 9 -o /tmp/${T}_c3 $C3 > gpurun_out/${T}_ncu_full_c3.log 2>&1
python - <<PY > gpurun_out/${T}_traffic_C3.json
import csv, json, subprocess
txt = subprocess.run(["ncu", "-i", "/tmp/${T}_c3.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, u = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}
out = {"workload": "configs[2] as named (10,000 entries)", "kernels": []}
tot = 0
for r in rows[2:]:
    g = lambda k: float(r[h.index(k)].replace(",", "")) * scale.get(u[h.index(k)], 1)
    k = {"kernel": r[h.index("Kernel Name")].split("(")[0], "ms": g("gpu__time_duration.sum") * 1e3,
         "dram_bytes": g("dram__bytes_read.sum") + g("dram__bytes_write.sum"),
         "issue_active": float(r[h.index("smsp__issue_active.avg.per_cycle_active")]),
         "warps_active_pct": float(r[h.index("sm__warps_active.avg.pct_of_peak_sustained_active")]),
         "inst_executed": float(r[h.index("smsp__inst_executed.sum")].replace(",", ""))}
    out["kernels"].append(k)
    if "crc" not in k["kernel"]:
        tot += k["dram_bytes"]
out["decode_phase_dram_bytes"] = tot
print(json.dumps(out, indent=1))
PY
rm -f /tmp/${T}_c3.ncu-rep
full() {  # name, kernel regex, command, launches to skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-2} -c 1 -o /tmp/${T}_prof_$1 $3 > gpurun_out/${T}_ncu_full_$1.log 2>&1
  python tools/ncu_summary.py /tmp/${T}_prof_$1.ncu-rep > gpurun_out/${T}_ncu_$1.json 2>/dev/null
  python tools/ncu_lines.py /tmp/${T}_prof_$1.ncu-rep 40 > gpurun_out/${T}_lines_$1.txt 2>/dev/null
  rm -f /tmp/${T}_prof_$1.ncu-rep
}
C1="python bench.py --workload c1 --entries 10000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
C3S="python bench.py --workload c3 --entries 2000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
full spec1 k_inflate_spec "$C1" 2
full lz k_inflate_lz "$C1" 2
full spec4 k_inflate_spec "$C3S" 2
full par k_inflate_lz "$C3S" 8   # (per pass: segment walk / PAR of three size groups, then the regular streams: launch 8 = PAR of the largest streams, second pass)
full translate k_seg_translate "$C3S" 3
C4Z="python bench.py --workload c4z --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
C5="python bench.py --workload c5 --entries 2048 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0"
full zseq k_zstd_seq "$C4Z" 1
full zlit k_zstd_lit "$C4Z" 1
full deflate k_deflate_chunks "$C5" 1
OTZ_ZSTD_TRACE=1 $C4Z 2>&1 > /dev/null | grep "otz zstd" | head -3 > gpurun_out/${T}_zstd_trace.txt
cuobjdump -sass otezip_b200/csrc/otz_shim.o 2>/dev/null | awk '/Function : .*k_seg_translate/{f=1} /Function : .*k_seg_window/{f=0} f' | grep -E "Function|UBLKCP|SYNCS|FENCE|LDS|STG|LDG" | head -60 > gpurun_out/${T}_sass_translate.txt
cuobjdump -sass otezip_b200/csrc/otz_shim.o 2>/dev/null | awk '/Function : .*k_crc_chunks/{f=1} /Function : .*k_crc_finalize/{f=0} f' | grep -E "Function|LDG|SHF|LOP3|SHFL" | head -80 > gpurun_out/${T}_sass_crc.txt
ls -la gpurun_out | grep ${T}_
