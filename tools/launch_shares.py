#!/usr/bin/env python
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[start]
kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[mu], 1e-6)
    name = r[kn].split('(')[0].split('::')[-1]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-32s launches %4d  %9.3f ms  %5.1f %%' % (k, v[0], v[1], 100 * v[1] / tot))
print('%-32s %24.3f ms' % ('total', tot))
