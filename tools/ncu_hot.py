"""Hot spots of an ncu source-page CSV (SASS): top instructions by stall samples with their dominant stall reason.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [top]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
si = h.index('# Samples')
stall = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
body = rows[hi + 1:]
tot = sum(int(r[si]) for r in body if len(r) == len(h))
print('total samples', tot)
order = sorted(range(len(body)), key=lambda k: -int(body[k][si]) if len(body[k]) == len(h) else 0)[:top]
for k in sorted(order):
    r = body[k]
    s = int(r[si])
    why = sorted(((int(r[i] or 0), n) for i, n in stall), reverse=True)[:2]
    print(f'{k:5d} {s:6d} {100 * s / tot:5.1f}%  {r[1].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}')
