#!/usr/bin/env python
"""profiles/traffic.json from tools/traffic_round.sh's CSV: DRAM bytes per bench step and stage (last step captured),
and the busiest issue pipe of each stage's kernel.  Usage: make_traffic_json.py traffic.csv "<workload name>" launches_per_step"""
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[start]
kn, mn, mv, idc = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
stage_of = dict(resize_level='pyramid', fast_cells='fast', quadtree_kernel='quadtree', describe_kernel='describe',
                match_pair_kernel='match', match_prepare='match', scc_merge='scc_merge')
launch = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= mv: continue
    d = launch.setdefault(r[idc], dict(name=r[kn]))
    d[r[mn]] = float(r[mv].replace(',', ''))
L = list(launch.values())
# the capture holds warm-up steps too: keep the last step = the last 1/(steps) share, detected by the first resize launch of the last group
firsts = [i for i, d in enumerate(L) if 'resize_level' in d['name'] and (i == 0 or 'resize_level' not in L[i - 1]['name'])]
# (bench.py ends with a one-image extraction for its candidate statistics: take the LARGEST group, not the last)
groups = [L[a:b] for a, b in zip(firsts, firsts[1:] + [len(L)])] if firsts else [L]
last = max(groups, key=lambda g: sum(d.get('gpu__time_duration.sum', 0) for d in g))
# the brute-force roofline runs (match_pair_kernel<.., 0>) follow the step: stop at the second match_prepare launch
cut = [i for i, d in enumerate(last) if 'match_prepare' in d['name']]
if len(cut) > 1: last = last[:cut[1]]
bytes_, pipes, times = collections.Counter(), {}, collections.Counter()
for d in last:
    st = next((v for k, v in stage_of.items() if k in d['name']), None)
    if not st: continue
    bytes_[st] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    t = d.get('gpu__time_duration.sum', 0)
    times[st] += t
    p = pipes.setdefault(st, collections.Counter())
    for k in ('alu', 'fma', 'lsu'):
        p[k] += t * d.get('sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active' % k, 0)
    p['issue'] += t * d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0)
out = dict(workload=sys.argv[2], source='ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (tools/traffic_round.sh), summed over the launches of one step',
           dram_bytes_per_step={k: v for k, v in bytes_.items()},
           pipe_pct={k: {q: round(v / times[k], 1) for q, v in p.items()} for k, p in pipes.items() if times[k]},
           ncu_ms_per_step={k: round(v / 1e6, 4) for k, v in times.items()})
json.dump(out, sys.stdout, indent=1)
