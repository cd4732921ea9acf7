#!/usr/bin/env python
"""Digest an ncu report: key metrics per launch and, with --hot N, the per-instruction issue counts of launch N.
    python tools/ncu_digest.py gpurun_out/x.ncu-rep [--hot 2 iters_per_launch]"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for i, h in enumerate(hdr):
    if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        vals = [r[i] for r in data]
        if "issue_stalled" in h and max(float(v or 0) for v in vals) < 0.15:
            continue
        print("%-90s %-8s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall:"), units[i], "  ".join(v[:38] for v in vals)))
if "--hot" in sys.argv:
    k = int(sys.argv[sys.argv.index("--hot") + 1])
    iters = float(sys.argv[sys.argv.index("--hot") + 2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    ia, isrc, iex, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    d = [r for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
    tot = sum(int(r[iex]) for r in d)
    print("instructions executed (source page): %d = %.1f per iteration" % (tot, tot / iters))
    base = int(d[0][ia], 16)
    cat = collections.Counter()
    out = []
    for r in d:
        e = int(r[iex]) / iters
        op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0]
        cat[op] += e
        if e > 0.02:
            out.append("%5x %7.2f %5s  %s" % (int(r[ia], 16) - base, e, r[ismp], r[isrc]))
    print([(a, round(b, 1)) for a, b in sorted(cat.items(), key=lambda x: -x[1])][:30])
    open("/tmp/hot.txt", "w").write("\n".join(out))
