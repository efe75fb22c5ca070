"""Top stall sites of an `ncu --page source --csv` export: python scripts/ncu_src_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
si = hdr.index('# Samples'); src = hdr.index('Source'); ex = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
print("total samples", tot)
order = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:n]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c]), hdr[c][6:]) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0), reverse=True)[:3]
    print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/tot:5.1f}%  x{r[ex]:>8s} {r[src].strip()[:70]:70s} {st}")
