#!/usr/bin/env python
"""Segment a kernel's SASS profile (ncu --page source, sass view) by execution count and print where samples go."""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
for k, r in enumerate(rows):
    if r and r[0] == 'Address':
        hdr = r; start = k + 1; break
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[start:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) >= len(hdr) - 5: data.append(r)
I = lambda r, h: int(r[idx[h]] or 0)
tot = sum(I(r, '# Samples') for r in data); totx = sum(I(r, 'Instructions Executed') for r in data)
print(len(data), 'sass instrs; samples', tot, 'executed', totx)
seg = []; cur = None
for k, r in enumerate(data):
    e = I(r, 'Instructions Executed')
    if cur is None or e != cur['e']:
        cur = dict(a=k, b=k, e=e, s=0, w=0, sh=0, lg=0, n=0, mio=0, math=0, ops=collections.Counter()); seg.append(cur)
    cur['b'] = k; cur['s'] += I(r, '# Samples'); cur['w'] += I(r, 'stall_wait'); cur['sh'] += I(r, 'stall_short_sb'); cur['lg'] += I(r, 'stall_long_sb'); cur['n'] += 1
    cur['mio'] += I(r, 'stall_mio'); cur['math'] += I(r, 'stall_math')
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[idx['Source']]); cur['ops'][m.group(2).split('.')[0] if m else '?'] += 1
base = min(s['e'] for s in seg if s['e'] > 0)
for s in seg:
    if s['s'] > 0.004 * tot or s['n'] > 40:
        print('i%5d-%5d n=%4d x%5.1f smp %5.1f%% (wait %4.1f short %4.1f long %4.1f mio %4.1f math %4.1f) instr %5.1f%%  %s' % (
            s['a'], s['b'], s['n'], s['e'] / base, 100 * s['s'] / tot, 100 * s['w'] / tot, 100 * s['sh'] / tot, 100 * s['lg'] / tot,
            100 * s['mio'] / tot, 100 * s['math'] / tot, 100 * s['n'] * s['e'] / totx, ' '.join('%s:%d' % kv for kv in s['ops'].most_common(6))))
