"""Aggregates an ncu launch list (gpu__time_duration.sum csv) of tests/cuda/transmil_time.py over one forward."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size')
data = rows[hi + 1:]
marker = sys.argv[2] if len(sys.argv) > 2 else '(4,391,1)'
starts = [i for i, r in enumerate(data) if r[gi].replace(' ', '') == marker]
seg = data[starts[1]:starts[2]]
agg = collections.OrderedDict(); tot = 0
for r in seg:
    key = (r[ki].split('(')[0].split('::')[-1][:40], r[gi])
    t = float(r[vi]) / 1e3; tot += t
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += t
for k, (c, t) in agg.items():
    print(f"{t:9.1f} us  x{c:3d}  {k[1]:>16}  {k[0]}")
print("launches", len(seg), "total us", round(tot, 1))
