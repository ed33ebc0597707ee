"""Kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean duration, share of the summed
kernel time.  usage: python tools/launches_summary.py launches.csv > summary.txt"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 10]
hdr = next(r for r in rows if 'Kernel Name' in r)
ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
iu = hdr.index('Metric Unit')
acc = defaultdict(list)
for r in rows:
    if r is hdr or r[im] != 'gpu__time_duration.sum':
        continue
    v = float(r[iv].replace(',', ''))
    unit = r[iu]
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1e-3)
    acc[r[ik]].append(v)
tot = sum(sum(v) for k, v in acc.items() if 'fp64_peak' not in k)
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    share = 0.0 if 'fp64_peak' in k else 100.0 * sum(v) / tot
    print('%5d launches  avg %9.2f us  share %5.1f%%  %s' % (len(v), sum(v) / len(v), share, k[:140]))
