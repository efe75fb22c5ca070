"""Summarise an ncu launch list (gpu__time_duration.sum per launch) over the LAST full step of bench.py."""
import csv, collections, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); mi = hdr.index('Metric Name')
names = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[1:] if r[mi] == 'gpu__time_duration.sum']
idx = [i for i, (n, _) in enumerate(names) if 'gather_rows' in n]
start, end = idx[-2], idx[-1]          # the last complete step
agg = collections.OrderedDict()
for n, v in names[start:end]:
    k = re.sub(r'\(.*', '', n).replace('void ', '').replace('svae::', '').replace('<unnamed>::', '')[:60]
    agg.setdefault(k, [0, 0]); agg[k][0] += v; agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{v/1e3:9.1f} us x{c:3d}  {100*v/tot:5.1f}%  {k}")
print(f"total {tot/1e3:.1f} us, {sum(v[1] for v in agg.values())} launches")
