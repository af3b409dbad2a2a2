#!/usr/bin/env python
"""Turn the artefacts a gpurun call left in gpurun_out/ into the tracked evidence under profiles/.

    python tools/refresh_profiles.py <full.ncu-rep> [round tag, default r2]
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def main(rep, tag="r2"):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    heads = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_head.py"), rep], capture_output=True, text=True).stdout
    lines_prof = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "env_compact.cu", "1.0"],
                                capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_k1_compact_65536envs.summary.txt"), "w").write(
        "ncu --set full --import-source on --clock-control none, one launch of each kernel of a steady-state env step\n"
        "(command: STEPS=40 PLAIN=1 python tools/step_timings.py, launches 105-107 = step t = 35)\n\n" + heads +
        "\nexecuted warp instructions / stall samples per source line of csrc/env_compact.cu (>= 1 % of the three kernels):\n" + lines_prof)
    rows = list(csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
            continue
        name = r[kn].split('(')[0][:70]
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none -k regex:compact_|env_place|env_alloc|env_feature|env_reward|env_step|env_reset|lines_from|gae_kernel|moments|standardize -c 500",
             "command: python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-spot-check   (every launch of this library's kernels: demand conversion, recording pass, warm-up, timed segment)",
             "(cold-cache, serialised per-launch times under the profiler: compare shares, not absolutes)", ""]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:15]:
        lines.append(f"{100 * v[1] / tot:6.2f}%  n={v[0]:4d}  avg={v[1] / v[0] / 1e6:9.3f} ms  {k}")
    open(os.path.join(P, f"{tag}_launch_list_summary.txt"), "w").write("\n".join(lines) + "\n")
    shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{tag}_launches.csv"))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def val(r, n):
        i = hdr.index(n)
        return float(r[i].replace(',', '')) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(units[i], 1)
    kernels, rd, wr = [], 0.0, 0.0
    for r in rows[2:]:
        a, b = val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
        kernels.append(dict(kernel=r[hdr.index('Kernel Name')].split('(')[0], dram_read=a, dram_write=b))
        rd, wr = rd + a, wr + b
    json.dump(dict(workload="large", envs=65536, layout="compact", kernel="compact split step: compact_place + compact_alloc + compact_feature kernels",
                   dram_bytes_per_launch=rd + wr, dram_read=rd, dram_write=wr, kernels=kernels,
                   source=f"profiles/{tag}_k1_compact_65536envs.summary.txt (ncu --set full, one launch of each kernel of one env step)"),
              open(os.path.join(P, "k1_traffic.json"), "w"), indent=1)
    for name in ("bench_default", "bench_wide", "bench_small", "bench_ippo", "bench_reference"):
        src = os.path.join(G, name + ".json")
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, f"{tag}_{name}.json"))
    for name in (f"{tag}_step_timings.log", f"{tag}_pipeline_timings.log", f"{tag}_pytest_gpu.log", f"{tag}_small_team_sweep.log"):
        if os.path.exists(os.path.join(G, name)):
            shutil.copy(os.path.join(G, name), os.path.join(P, name))
    print("\n".join(lines))
    print(f"DRAM traffic per env-step: {(rd + wr) / 65536:.0f} B (read {rd / 65536:.0f}, write {wr / 65536:.0f})")


if __name__ == "__main__":
    main(sys.argv[1], *(sys.argv[2:3]))
