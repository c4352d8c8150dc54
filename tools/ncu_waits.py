"""Where the roles of a warp-specialised kernel wait: every mbarrier try-wait site of an ncu report's SASS page with its
share of the warp-stall samples and its execution count (retries), per kernel.
    ncu -i X.ncu-rep --page source --csv > f.csv;  python tools/ncu_waits.py f.csv [kernel indices...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        ks.append(cur)
    elif r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and r:
        cur['rows'].append(r)
want = [int(x) for x in sys.argv[2:]] or range(len(ks))
for kk in want:
    k = ks[kk]
    h = k['hdr']
    si, ie = h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
    tot = sum(float(r[si]) for r in k['rows']) or 1.0
    mma = sum(float(r[ie]) for r in k['rows'] if 'UTCHMMA' in r[1] or 'UTCQMMA' in r[1])
    print('== launch %d  %s\n   warp-stall samples %d, tcgen05.mma issued %d' % (kk, k['name'][:90], tot, mma))
    for i, r in enumerate(k['rows']):
        ins = r[1].strip()
        if 'SYNCS' in ins and 'TRYWAIT' in ins:
            s = float(r[si]) + float(k['rows'][i + 1][si])
            print('   sass %5d  %5.1f%% of samples  executed %9s x  %s' % (i, 100 * s / tot, r[ie], ins[:70]))
