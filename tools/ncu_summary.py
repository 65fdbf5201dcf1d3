#!/usr/bin/env python
"""One-paragraph summary of an .ncu-rep (first profiled launch): duration, DRAM bytes, pipe / issue utilisation,
stall reasons.  usage: tools/ncu_summary.py file.ncu-rep [algorithmic_bytes]"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
algo = float(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
g = lambda k: (float(v[h.index(k)].replace(",", "")), u[h.index(k)]) if k in h else (None, "")
name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}
t, tu = g("gpu__time_duration.sum")
rd, ru = g("dram__bytes_read.sum")
wr, wu = g("dram__bytes_write.sum")
t_s = t * scale.get(tu, 1)
traffic = rd * scale.get(ru, 1) + wr * scale.get(wu, 1)
out = {"kernel": name.split("(")[0], "duration_ms_under_ncu": t_s * 1e3, "dram_read_bytes": rd * scale.get(ru, 1),
       "dram_write_bytes": wr * scale.get(wu, 1), "dram_traffic_bytes": traffic}
if algo:
    out["algorithmic_bytes"] = algo
    out["traffic_over_algorithmic"] = traffic / algo
for k in ["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]:
    x, _ = g(k)
    if x is not None:
        out[k] = x
st = {h[i].replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v[i]) for i in range(len(h))
      if "pcsamp_warps_issue_stalled" in h[i] and "not_issued" not in h[i]}
tot = sum(st.values()) or 1
out["stall_share_pct"] = {k: round(100 * x / tot, 1) for k, x in sorted(st.items(), key=lambda kv: -kv[1])[:7]}
print(json.dumps(out, indent=1))
