#!/usr/bin/env python
"""Summarise an ncu report of the env-step kernels: per kernel the headline metrics and stall reasons, then the
executed instructions / stall samples per source region and the hottest source lines (needs -lineinfo and
--import-source on).

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
# (label, text that opens the region in env_core.cuh); a line belongs to the last marker at or above it
MARKS = [("helpers (arithmetic, team votes, tables)", None),
         ("write_obs_pipeline", "MDEV void write_obs_pipeline"),
         ("write_obs_row", "MDEV void write_obs_row"),
         ("allocate_orders: setup + staging", "MDEV void allocate_orders"),
         ("allocate_orders: lane masks", "if constexpr (LaneAlloc<G, CAPS>::value) {"),
         ("allocate_orders: lane chains", "int rem[2] = {0, 0}"),
         ("allocate_orders: lost-order flags", "if (s_sreg[j] & 0x8000) smem_add"),
         ("allocate_orders: dense (narrow teams / generic)", "      for (int j = 0; j < cn; ++j) {"),
         ("step_env (fused kernel phases 1, 3, 4)", "MDEV void step_env"),
         ("reset_env", "MDEV void reset_env")]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def regions():
    src = open(os.path.join(ROOT, "marl-sc_b200", "csrc", "env_core.cuh")).read().split("\n")
    out = []
    for label, pat in MARKS:
        if pat is None:
            out.append((label, 1))
            continue
        hits = [i + 1 for i, l in enumerate(src) if pat in l]
        if hits:
            out.append((label, hits[0]))
    return out


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
    marks = regions()

    def region(f, n):
        if f != "env_core.cuh":
            return f
        name = marks[0][0]
        for label, at in marks:
            if n >= at:
                name = label
        return name

    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]))))
    kern, cur, hdr, last = None, None, None, None
    per = collections.OrderedDict()        # kernel -> {(file, line): [inst, samples, sass, text]}
    seen = set()
    for r in rows:
        if len(r) >= 2 and r[0] == "Function Name":
            kern = r[1].replace("marlsc::", "").split("(")[0].replace("(int)", "")
            per.setdefault(kern, collections.defaultdict(lambda: [0, 0, 0, '']))
            continue
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split('/')[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            iex, ism = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr and len(r) == len(hdr) and kern:
            if r[0]:
                last = (cur, int(r[0]))
                per[kern][last][3] = r[1]
                continue
            try:
                key = (kern, int(r[2], 16))
                ex, sm = int(r[iex] or 0), int(r[ism] or 0)
            except ValueError:
                continue
            if key in seen or last is None:      # the same SASS row is listed under every file it is attributed to
                continue
            seen.add(key)
            a = per[kern][last]
            a[0] += ex
            a[1] += sm
            a[2] += 1
    for kern, agg in per.items():
        tot = sum(v[0] for v in agg.values()) or 1
        ts = sum(v[1] for v in agg.values()) or 1
        print(f"\n== {kern}: {tot} warp instructions, {ts} stall samples, {sum(v[2] for v in agg.values())} SASS instructions")
        reg = collections.defaultdict(lambda: [0, 0])
        for (f, n), v in agg.items():
            k = region(f, n)
            reg[k][0] += v[0]
            reg[k][1] += v[1]
        for k, v in sorted(reg.items(), key=lambda kv: -kv[1][0]):
            if v[0] * 200 > tot or v[1] * 200 > ts:
                print(f"  {k:52s} {100 * v[0] / tot:5.1f}% inst  {100 * v[1] / ts:5.1f}% samples")
        print("  top source lines by stall samples:")
        for (f, n), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
            print(f"    {100 * v[0] / tot:5.1f}% inst {100 * v[1] / ts:5.1f}% smp sass={v[2]:4d} {f}:{n:5d} {v[3].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1])
