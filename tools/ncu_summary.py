#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep as JSON (the files under profiles/ncu_*.json).
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep "how it was captured" [kernel index] > profiles/ncu_xxx.json"""
import csv, json, subprocess, sys

KEEP = ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")

rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)]
d = {h: {"value": v, "unit": u} for h, u, v in zip(hdr, units, vals)}
res = {"source": note, "kernel": d.get("Kernel Name", {}).get("value"), "metrics": {}, "stall_per_issue": {}}
for k in KEEP:
    if k in d:
        res["metrics"][k] = d[k]
for k, v in d.items():
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        res["stall_per_issue"][k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v["value"])
json.dump(res, sys.stdout, indent=1)
print()
