#!/usr/bin/env python
"""Per-layer table from an ncu --csv metrics log of tools/bench_conv.py (three precisions, 10 conv launches each)."""
import csv, re, collections, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
byid = collections.OrderedDict()
for r in csv.DictReader(lines):
    byid.setdefault(r['ID'], {'name': re.sub(r'\(.*', '', r['Kernel Name'])})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
L = list(byid.values())
per = 10
calls = len(L) // (3 * per)
names = ['conv2', 'conv5', 'conv7', 'conv10', 'conv12', 'conv14', 'dec1.1', 'dec1.3', 'dec2.1', 'dec2.3']
for pi, prec in enumerate(['f16x3_1cta', 'f16x3 (pair)', 'f16']):
    base = (pi * calls + calls - 1) * per
    tot = 0
    print(prec)
    for j in range(per):
        r = L[base + j]
        d = r['gpu__time_duration.sum']
        d = d / 1e6 if d > 1e4 else d
        tot += d
        print('  %-7s %-36s %7.3f ms  tensor %5.1f%%  clk %.2f GHz  lts %6.2f GB' % (names[j], r['name'][-36:], d,
              r['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], r['sm__cycles_elapsed.max'] / (d * 1e-3) / 1e9, r['lts__t_bytes.sum'] / 1e9))
    print('  total %.3f ms' % tot)
