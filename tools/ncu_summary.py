#!/usr/bin/env python
"""Key issue/stall metrics of one ncu capture (first kernel in the report). usage: ncu_summary.py file.ncu-rep [races]"""
import csv, subprocess, sys
rep = sys.argv[1]; races = float(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, v = rows[0], rows[2]
d = dict(zip(h, v))
def g(k):
    return float(d[k].replace(",", "")) if k in d and d[k] not in ("", "n/a") else float("nan")
inst = g("smsp__inst_executed.sum")
print("duration ms", g("gpu__time_duration.sum"), " regs", d.get("launch__registers_per_thread"))
print("inst executed", inst, " per race", inst / races if races else "")
for k in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "smsp__sass_average_branch_targets_threads_uniform.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"):
    print(f"{k:75s} {d.get(k)}")
st = sorted(((float(x.replace(",", "")), k) for k, x in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and x not in ("", "n/a")), reverse=True)
for x, k in st[:9]:
    print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:25s} {x:.2f}")
