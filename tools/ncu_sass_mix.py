"""Opcode mix and stall hot spots of one kernel of an .ncu-rep (source page).
    python tools/ncu_sass_mix.py REP KERNEL_REGEX [launch_skip]"""
import csv, io, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else '0'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre, '--launch-skip', skip,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:100])
h = rows[1]
iS, iE, iW = h.index('Source'), h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
mix, stall = collections.Counter(), collections.Counter()
tot = 0
lines = []
for r in rows[2:]:
    try:
        n, w = int(r[iE]), int(r[iW])
    except Exception:
        continue
    op = r[iS].split()
    op = op[1] if op and op[0].startswith('@') else (op[0] if op else '?')
    op = op.split('.')[0]
    mix[op] += n; stall[op] += w; tot += n
    lines.append((w, n, r[iS].strip()))
print('total warp instructions', tot)
ws = sum(stall.values()) or 1
for op, n in mix.most_common(28):
    print('  %-10s %12d %5.1f%%   stall %5.1f%%' % (op, n, 100.0 * n / tot, 100.0 * stall[op] / ws))
print('top stalled instructions:')
for w, n, s in sorted(lines, reverse=True)[:int(sys.argv[4]) if len(sys.argv) > 4 else 25]:
    print('  %6d %5.1f%%  x%-9d %s' % (w, 100.0 * w / ws, n, s[:100]))
