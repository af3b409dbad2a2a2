#!/usr/bin/env python
"""Headline counters of every kernel in an ncu report (time, DRAM bytes, occupancy limits, issue utilisation, stalls).

    python tools/ncu_head.py report.ncu-rep
"""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_red.sum']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[h.index('Kernel Name')][:110])
    for w in WANT:
        if w in h:
            print(f"  {w:84s} {r[h.index(w)]:>18s} {u[h.index(w)]}")
    for i, n in enumerate(h):
        if 'warps_issue_stalled' in n and n.endswith('per_issue_active.ratio'):
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if v > 0.3:
                print(f"  {n:84s} {r[i]:>18s}")
