#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + top source lines) into a small text file for profiles/."""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_mio_throttle']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = ['# ' + rep.split('/')[-1], note, '']
    for name in ['Kernel Name']:
        if name in hdr:
            lines.append('kernel: ' + vals[hdr.index(name)])
    for i, h in enumerate(hdr):
        if h in WANT:
            lines.append('%-75s %s %s' % (h, vals[i], units[i]))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    cur, per, hdr2 = None, {}, None
    for r in csv.reader(src.splitlines()):
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr2 = r
            continue
        if hdr2 and len(r) > 7 and r[0].isdigit() and r[2] == '-':
            try:
                per[(cur, int(r[0]))] = (int(r[6]), int(r[7]), r[1][:110])
            except ValueError:
                pass
    ts = sum(v[0] for v in per.values()) or 1
    te = sum(v[1] for v in per.values()) or 1
    byfile = collections.Counter()
    for (f, _), v in per.items():
        byfile[f] += v[0]
    lines += ['', 'warp-stall samples by file: ' + ', '.join('%s %.1f%%' % (f, 100 * c / ts) for f, c in byfile.most_common(5)),
              '', 'hottest source lines (share of stall samples / of executed instructions):']
    for (f, ln), (s, e, text) in sorted(per.items(), key=lambda kv: -kv[1][0])[:18]:
        lines.append('%5.1f%% %5.1f%%  %s:%d  %s' % (100 * s / ts, 100 * e / te, f, ln, text.strip()))
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('wrote', out)


if __name__ == '__main__':
    main()
