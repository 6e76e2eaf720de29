#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.
usage: ncu_lines.py prof_cs.csv [top_n]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = defaultdict(lambda: defaultdict(float))
src_text = {}
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] in ('File Name', 'File Path'):
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    line = r[0]
    key = (cur_file, line)
    if r[1].strip():
        src_text[key] = r[1].strip()
    d = dict(zip(hdr[4:], r[4:]))

    def num(c):
        try:
            return float(d.get(c, 0) or 0)
        except ValueError:
            return 0.
    a = agg[key]
    a['samples'] += num('# Samples')
    a['inst'] += num('Instructions Executed')
    for s in ('stall_barrier', 'stall_long_sb', 'stall_wait', 'stall_math', 'stall_short_sb', 'stall_not_selected', 'stall_selected',
              'stall_branch_resolving', 'stall_lg', 'stall_dispatch', 'stall_no_inst', 'stall_mio'):
        a[s] += num(s)
    a['local'] += num('L2 Theoretical Sectors Local')
    a['shared_wf'] += num('L1 Wavefronts Shared')
tot = sum(a['samples'] for a in agg.values())
toti = sum(a['inst'] for a in agg.values())
print('total samples %.0f, warp instructions %.3e' % (tot, toti))
stalls = defaultdict(float)
for a in agg.values():
    for k, v in a.items():
        if k.startswith('stall_'):
            stalls[k] += v
print('stall mix: ' + ', '.join('%s %.1f%%' % (k[6:], 100 * v / tot) for k, v in sorted(stalls.items(), key=lambda x: -x[1])))
print('%-18s %6s %7s %7s %6s %6s %6s %6s  %s' % ('file:line', 'smpl%', 'inst%', 'local%', 'bar', 'lsb', 'wait', 'math', 'source'))
totl = sum(a['local'] for a in agg.values()) or 1
for key, a in sorted(agg.items(), key=lambda x: -x[1]['samples'])[:top_n]:
    print('%-18s %6.2f %7.2f %7.2f %6.0f %6.0f %6.0f %6.0f  %s' % (
        '%s:%s' % key, 100 * a['samples'] / tot, 100 * a['inst'] / toti, 100 * a['local'] / totl, a['stall_barrier'], a['stall_long_sb'],
        a['stall_wait'], a['stall_math'], src_text.get(key, '')[:90]))
