"""Top stalled SASS instructions of an ncu report's source page:  ncu -i X.ncu-rep --page source --csv > f; python tools/ncu_hot.py f [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
si = hdr.index('Warp Stall Sampling (All Samples)')
lsb = hdr.index('stall_long_sb')
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((float(r[si]), i, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print('total samples', tot)
for v, i, r in sorted(data, key=lambda x: -x[0])[:n]:
    print('%7.0f %5.1f%% line %5d  long_sb %6s  %s' % (v, 100 * v / tot, i, r[lsb], r[1].strip()[:110]))
