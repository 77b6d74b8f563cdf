import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; ci = {n: i for i, n in enumerate(h)}
body = rows[hi + 1:]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
body = body[int(len(body) * frac):]
agg = defaultdict(lambda: [0, 0.0])
for r in body:
    name = r[ci['Kernel Name']][:70]; v = float(r[ci['Metric Value']].replace(',', '')); u = r[ci['Metric Unit']]
    v = v * {'us': 1, 'ms': 1000, 'ns': 1e-3, 's': 1e6}.get(u, 1)
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:72s} {v[0]:4d} {v[1]:10.1f} us {100 * v[1] / tot:5.1f}%")
print("total us", tot)
