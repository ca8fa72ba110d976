"""Last-step launch list from an ncu --metrics gpu__time_duration.sum CSV: python tools/launch_summary.py file.csv N"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 13
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
data = [(r[kn].split('(')[0], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv]
last = data[-n:]
tot = sum(v for _, v in last)
for k, v in last:
    print(f'{v / 1e3:9.1f} us {100 * v / tot:5.1f}%  {k}')
print(f'total {tot / 1e3:.1f} us over {n} launches (ncu per-launch times are cold-cache and serialised: compare shares)')
