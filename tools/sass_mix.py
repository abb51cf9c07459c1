#!/usr/bin/env python
"""Opcode mix of one kernel from `ncu -i rep --page source --csv --print-source sass` (stdin or file):
warp-level instructions executed and stall samples per opcode, plus the hottest SASS lines."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr = rows[1]
iS, iI, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
mix, smp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iI: continue
    src = r[iS].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    n = int(r[iI] or 0); s = int(r[iN] or 0)
    mix[op] += n; smp[op] += s; tot += n
ts = sum(smp.values()) or 1
print('total warp instructions %d, samples %d' % (tot, ts))
for op, n in mix.most_common(28):
    print('  %-28s %6.2f %% inst  %6.2f %% samples' % (op, 100.0 * n / tot, 100.0 * smp[op] / ts))
