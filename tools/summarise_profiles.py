#!/usr/bin/env python
"""Turn the raw ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarise_profiles.py r01

Reads gpurun_out/launches_<tag>.csv (per-launch gpu__time_duration) and gpurun_out/{conv,lm}_<tag>.ncu-rep
(ncu --set full) and writes profiles/<tag>_launches.md, profiles/<tag>_{conv,lm}_full.csv.
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.max", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]

f = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(f):
    lines = [l for l in open(f) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        agg.setdefault(name, []).append(float(r["Metric Value"].replace(",", "")) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, "%s_launches.md" % tag), "w") as o:
        o.write("# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 3`\n\n" % tag)
        o.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes. %d launches.\n\n" % len(rows))
        o.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            o.write("| `%s` | %d | %.1f | %.1f%% | %.1f |\n" % (k[:70], len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v)))
    print("wrote launches summary")

for kind in ("conv", "lm", "lmstress"):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (kind, tag))
    raw = os.path.join(G, "%s_%s_raw.csv" % (kind, tag))       # exported on the GPU box by tools/gpu_session.sh
    if os.path.exists(raw):
        out = open(raw).read()
    elif os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    else:
        continue
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(os.path.join(P, "%s_%s_full.csv" % (tag, kind)), "w") as o:
        w = csv.writer(o)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(rows) - 2)])
        for i in idx:
            w.writerow([hdr[i], units[i]] + [r[i] for r in rows[2:]])
    print("wrote", kind)
