#!/usr/bin/env python
"""Executed warp instructions and stall samples per source line of one file, in source order, from an ncu report
(needs -lineinfo and --import-source on).

    python tools/ncu_lines.py report.ncu-rep env_alloc.cuh [min percent, default 0.3]
"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, want, minpct=0.3):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    cur, hdr, last = None, None, None
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, ''])     # (file, line) -> inst, samples, sass, thread inst, text
    seen = set()
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split('/')[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            iex, ism, ith = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Thread Instructions Executed")
            continue
        if hdr and len(r) == len(hdr):
            if r[0]:
                last = (cur, int(r[0]))
                agg[last][4] = r[1]
                continue
            try:
                key = int(r[2], 16)
                ex, sm, th = int(r[iex] or 0), int(r[ism] or 0), int(r[ith] or 0)
            except ValueError:
                continue
            if key in seen or last is None:
                continue
            seen.add(key)
            a = agg[last]
            a[0] += ex; a[1] += sm; a[2] += 1; a[3] += th
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"total {tot} warp instructions, {ts} samples")
    byfile = collections.defaultdict(lambda: [0, 0])
    for (f, n), v in agg.items():
        byfile[f][0] += v[0]; byfile[f][1] += v[1]
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:40s} {100 * v[0] / tot:5.1f}% inst {100 * v[1] / ts:5.1f}% smp")
    for (f, n), v in sorted(agg.items()):
        if f == want and (100 * v[0] / tot >= minpct or 100 * v[1] / ts >= minpct):
            print(f"{n:5d} {100 * v[0] / tot:5.1f}%i {100 * v[1] / ts:5.1f}%s sass={v[2]:3d} thr={v[3] / max(v[0], 1):4.1f} | {v[4].rstrip()[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.3)
