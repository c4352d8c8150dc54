"""Condense an .ncu-rep (ncu --set full) into the text summary committed under profiles/.
    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/r01_X.txt ["note"]"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__average_warp_latency_per_inst_issued.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
]


def page(rep, which):
    out = subprocess.run(['ncu', '-i', rep, '--page', which, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    raw = page(rep, 'raw')
    hdr, units = raw[0], raw[1]
    lines = ['# ncu --set full --clock-control none  (%s)' % rep.split('/')[-1]]
    if note:
        lines.append('# ' + note)
    name_i = hdr.index('Kernel Name')
    for vals in raw[2:]:
        lines.append('')
        lines.append('kernel: ' + vals[name_i])
        d = dict(zip(hdr, zip(units, vals)))
        for k in KEYS:
            if k in d:
                lines.append('  %-88s %12s %s' % (k, d[k][1], d[k][0]))
    src = page(rep, 'source')
    if len(src) > 2:
        h = src[1]
        if 'Warp Stall Sampling (All Samples)' in h:
            si = h.index('Warp Stall Sampling (All Samples)')
            rows = []
            for r in src[2:]:
                try:
                    rows.append((float(r[si]), r[1].strip()))
                except Exception:
                    pass
            tot = sum(v for v, _ in rows) or 1.0
            lines.append('')
            lines.append('top stalled SASS instructions (warp stall samples, % of all samples; first kernel of the report):')
            for v, ins in sorted(rows, key=lambda x: -x[0])[:14]:
                lines.append('  %6.0f %5.1f%%  %s' % (v, 100 * v / tot, ins[:110]))
    open(dst, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines[:40]))


if __name__ == '__main__':
    main()
