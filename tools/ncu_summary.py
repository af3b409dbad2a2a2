#!/usr/bin/env python
"""Summarise an ncu report of the env-step kernel: headline metrics, stall reasons and executed
instructions / stall samples per kernel phase (needs -lineinfo and --import-source on).

    python tools/ncu_summary.py gpurun_out/foo.ncu-rep [> profiles/foo.summary.txt]
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index('Kernel Name')][:100])
        for w in WANT:
            if w in hdr:
                print(f"  {w:90s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    if float(r[i]) > 0.3:
                        print(f"  {h:90s} {r[i]:>18s}")
                except ValueError:
                    pass
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]))))
    cur, hdr, last = None, None, None
    agg = collections.defaultdict(lambda: [0, 0, 0, ''])
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split('/')[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            iex, ism = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr and len(r) == len(hdr):
            try:
                ex, sm = int(r[iex] or 0), int(r[ism] or 0)
            except ValueError:
                continue
            k = (cur, r[0])
            if r[0]:
                last = k
                agg[k][3] = r[1].strip()[:80]
            else:
                k = last
            agg[k][0] += ex
            agg[k][1] += sm
            agg[k][2] += 1
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    src = open(os.path.join(ROOT, "marl-sc_b200", "csrc", "env_core.cuh")).read().splitlines()

    def find(pat):
        return next(i + 1 for i, l in enumerate(src) if pat in l)
    marks = [(find("MDEV void write_obs_row"), "helpers"), (find("MDEV void step_env"), "write_obs_row"),
             (find("---- phase 2"), "phase1 (orders/arrivals)"), (find("for (int j = 0; j < cn; ++j) {"), "order staging"),
             (find("---- phase 3"), "phase2 (allocation)"), (find("---- phase 4"), "phase3 (features/costs)"),
             (10 ** 9, "phase4 (rewards)")]

    def phase(k):
        f, l = k
        if f != 'env_core.cuh':
            return f or "?"
        try:
            l = int(l)
        except ValueError:
            return "?"
        for lim, name in marks:
            if l < lim:
                return name
    ph = collections.defaultdict(lambda: [0, 0])
    for k, v in agg.items():
        ph[phase(k)][0] += v[0]
        ph[phase(k)][1] += v[1]
    print(f"per phase (executed warp instructions {tot}, stall samples {ts}):")
    for k, v in sorted(ph.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:28s} {100 * v[0] / tot:5.1f}% inst {100 * v[1] / ts:5.1f}% samples")
    print("top source lines by stall samples:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        print(f"  {100 * v[0] / tot:5.1f}% inst {100 * v[1] / ts:5.1f}% smp sass={v[2]:4d} {k[0]}:{k[1]:>4s} {v[3]}")


if __name__ == "__main__":
    main(sys.argv[1])
