"""List the SASS instructions of an `ncu --page source --csv` dump that collect the most stall samples, with their dominant stall
reason.  usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME | python profiles/hot_lines.py [N]"""
import csv, sys
n_top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for k, hi in enumerate(his):
    end = his[k + 1] if k + 1 < len(his) else len(rows)
    h = rows[hi]; col = {n: i for i, n in enumerate(h)}
    stall_cols = [n for n in h if n.startswith('stall_')]
    data = [r for r in rows[hi + 1:end] if len(r) >= len(h) and r[0] != 'Address']
    f = lambda r, c: float(r[col[c]] or 0)
    tot = sum(f(r, '# Samples') for r in data) or 1
    print('kernel %d: total samples %.0f, %d instructions, executed warp-instructions %.0f' % (k, tot, len(data), sum(f(r, 'Instructions Executed') for r in data)))
    order = sorted(range(len(data)), key=lambda i: -f(data[i], '# Samples'))[:n_top]
    for i in sorted(order):
        r = data[i]
        st = max(stall_cols, key=lambda s: f(r, s))
        print('%5d %6.0f %5.1f%%  %-14s %-80s' % (i, f(r, '# Samples'), 100 * f(r, '# Samples') / tot, st[6:], r[col['Source']][:80]))
