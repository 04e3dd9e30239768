#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the few numbers we track; usage: tools_ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
out = []
for v in rows[2:]:
    d = {h[i]: (v[i], u[i]) for i in range(len(h))}
    keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
            'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
            'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
    keys += sorted(k for k in h if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio'))
    for k in keys:
        if k in d:
            out.append("%s = %s %s" % (k, d[k][0], d[k][1]))
    out.append("")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
