"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    n = r[kn].split('(')[0]
    v = float(r[mv].replace(',', ''))
    u = r[mu]
    v = v / 1e3 if u in ('ns', 'nsecond') else v if u in ('us', 'usecond') else v * 1e3 if u in ('ms', 'msecond') else v
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
for n, (c, v) in agg.items():
    print(f"{n[:60]:60s} {c:5d} {v / 1e3:9.3f} ms  {v / c:8.1f} us each  {100 * v / tot:5.1f} %")
print(f"total {tot / 1e3:.3f} ms")
