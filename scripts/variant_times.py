"""Per-kernel launch times of a tuning variant from an ncu launch list (gpu__time_duration.sum):
python scripts/variant_times.py tag file.csv -- launches shorter than half the longest of their kernel (early exits) are left out."""
import csv
import sys
from collections import defaultdict

tag, path = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
header = next(r for r in rows if 'Kernel Name' in r)
kn, mv, mu = header.index('Kernel Name'), header.index('Metric Value'), header.index('Metric Unit')
times = defaultdict(list)
for r in rows:
    if r is header or len(r) <= mv or r[kn] == 'Kernel Name':
        continue
    try:
        v = float(r[mv].replace(',', ''))
    except ValueError:
        continue
    unit = r[mu]
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'msecond': 1e3, 'usecond': 1.0, 'nsecond': 1e-3}.get(unit, 1.0)
    name = r[kn].split('(')[0][:60]
    times[name].append(v)
for name, ts in sorted(times.items()):
    big = [t for t in ts if t > 0.5*max(ts)]
    print('%-8s %-62s n=%3d avg %9.1f us  min %9.1f' % (tag, name, len(big), sum(big)/len(big), min(big)))
