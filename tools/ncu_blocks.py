#!/usr/bin/env python
"""Aggregate warp-stall samples of an .ncu-rep source page into basic-block-like groups (split at barrier / TMEM / MMA ops)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[start]; ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[start + 1:] if len(r) == len(h)]
stall = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
blk = []; cur = None
def flush():
    global cur
    if cur and cur['samples'] > 0: blk.append(cur)
    cur = None
for r in data:
    src = r[ix['Source']]
    toks = [t for t in src.split() if not t.startswith('@')]
    op = toks[0] if toks else '?'
    n = int(r[ix['# Samples']] or 0); ex = int(r[ix['Instructions Executed']] or 0)
    key = any(k in src for k in ('SYNCS', 'LDTM', 'STTM', 'UTCHMMA', 'UTCBAR', 'BAR.', 'EXIT', 'UTMALDG', 'NANOSLEEP'))
    if key:
        flush()
        blk.append({'first': r[ix['Address']][-5:], 'ops': {op: 1}, 'samples': n, 'exec': ex, 'n': 1,
                    'st': {s[6:]: int(r[ix[s]] or 0) for s in stall}, 'key': src[:60]})
        continue
    if cur is None or abs(ex - cur['exec']) > 0.2 * max(ex, cur['exec'], 1):
        flush(); cur = {'first': r[ix['Address']][-5:], 'ops': {}, 'samples': 0, 'exec': ex, 'n': 0, 'st': {s[6:]: 0 for s in stall}, 'key': None}
    cur['ops'][op] = cur['ops'].get(op, 0) + 1; cur['samples'] += n; cur['n'] += 1
    for s in stall: cur['st'][s[6:]] += int(r[ix[s]] or 0)
flush()
tot = sum(b['samples'] for b in blk)
print("total samples", tot)
for b in blk:
    if b['samples'] < thr * tot: continue
    st = {k: v for k, v in b['st'].items() if v > 0.08 * b['samples']}
    ops = dict(sorted(b['ops'].items(), key=lambda kv: -kv[1])[:6])
    print(b['first'], 'n=%d exec=%d samples=%d (%.1f%%)' % (b['n'], b['exec'], b['samples'], 100 * b['samples'] / tot), b['key'] or ops, st)
