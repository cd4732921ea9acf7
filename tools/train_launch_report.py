"""Per-kernel totals of the LAST step in an ncu launch list of `tools/bench_train.py B 1 native` (time, tensor-pipe activity,
DRAM bytes)."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ik, im, iv, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = defaultdict(dict)
for r in rows[1:]:
    per[(r[ii], r[ik])][r[im]] = float(r[iv].replace(",", ""))
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
ids = sorted(int(i) for i, _ in per)
cut = ids[len(ids) * 2 // 3] if ids else 0           # the run is 2 warm-up steps + 1 measured step: keep the last third
for (i, k), m in per.items():
    if int(i) < cut:
        continue
    name = k.split("(")[0].replace("void ", "").replace("ha::", "")[:48]
    a = agg[name]
    t = m.get("gpu__time_duration.sum", 0.0)
    a[0] += 1; a[1] += t; a[2] += t * m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a[3] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | ms | share | tensor-pipe active (time-weighted) | DRAM GB |\n|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.3f | %.1f%% | %.1f%% | %.2f |" % (name, a[0], a[1] / 1e6, 100 * a[1] / tot, a[2] / max(a[1], 1e-9), a[3] / 1e9))
print("total %.3f ms over %d launches" % (tot / 1e6, sum(a[0] for a in agg.values())))
