"""Split an ncu source page of the single-pass kernel at its CTA barriers and print, per phase, executed warp-instructions per
pixel, the share of stall samples (a proxy for time) and the top opcodes / stall reasons.
usage: ncu -i X.ncu-rep --page source --csv | python profiles/phase_counts.py [pixels]"""
import csv, sys, collections
pixels = float(sys.argv[1]) if len(sys.argv) > 1 else 2263040.0
rows = list(csv.reader(sys.stdin))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]
rows = rows[:his[1]] if len(his) > 1 else rows      # `--kernel-name` exports repeat the table: keep the first copy
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
phases, cur = [], dict(inst=0, samples=0, ops=collections.Counter(), stalls=collections.Counter())
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    src = r[col['Source']].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op = op.split('.')[0]
    n = float(r[col['Instructions Executed']] or 0)
    cur['inst'] += n; cur['samples'] += float(r[col['# Samples']] or 0); cur['ops'][op] += n
    for s in stall_cols: cur['stalls'][s[6:]] += float(r[col[s]] or 0)
    if op == 'BAR':
        phases.append(cur); cur = dict(inst=0, samples=0, ops=collections.Counter(), stalls=collections.Counter())
phases.append(cur)
ti = sum(p['inst'] for p in phases); ts = sum(p['samples'] for p in phases)
print('total warp-instructions per pixel-warp: %.0f' % (ti / (pixels / 32)))
for k, p in enumerate(phases):
    tot = sum(p['stalls'].values()) or 1
    print('segment %d: %6.0f inst/px (%4.1f%%)  time share %4.1f%%' % (k, p['inst'] / (pixels / 32), 100 * p['inst'] / ti, 100 * p['samples'] / ts))
    print('    opcodes: ' + ', '.join('%s %.0f' % (o, n / (pixels / 32)) for o, n in p['ops'].most_common(12)))
    print('    stalls:  ' + ', '.join('%s %.0f%%' % (s, 100 * n / tot) for s, n in p['stalls'].most_common(7)))
