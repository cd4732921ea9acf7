#!/usr/bin/env python
"""From a SASS source-page CSV (tools/gpu_session.sh export_rep): every mbarrier try-wait with how often it executed —
first tries vs spins tells which barrier a warp-specialised kernel actually blocks on."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Address' in r]
h = rows[hi[0]]
isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
end = hi[1] if len(hi) > 1 else len(rows)
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows[:hi[0]]:
    if r and 'Kernel Name' in r[0]: print(r[1][:90])
sass = [r for r in rows[hi[0] + 1:end] if len(r) > iex]
for k, r in enumerate(sass):
    s = r[isrc]
    if 'TRYWAIT' in s and num(r[iex]) > 0:
        print('%5d samples %7d executed %10s  %s' % (k, num(r[isamp]), r[iex], s.strip()[:70]))
