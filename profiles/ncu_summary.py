#!/usr/bin/env python
"""Text summary of an .ncu-rep (one block per captured launch): duration, DRAM bytes, occupancy, issue rate, stall mix.
usage: python profiles/ncu_summary.py report.ncu-rep [n_top_sass_lines]"""
import csv, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 6
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio']
kn = h.index('Kernel Name') if 'Kernel Name' in h else None
for v in rows[2:]:
    if not v or len(v) < len(h) // 2:
        continue
    print('### ' + (v[kn][:150] if kn is not None else 'kernel'))
    for w in want:
        if w in h:
            i = h.index(w); print('  %-68s %s %s' % (w, v[i], rows[1][i]))
    st = []
    for i, name in enumerate(h):
        if 'issue_stalled' in name and name.endswith('per_issue_active.ratio'):
            try: st.append((float(v[i]), name.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError: pass
    print('  stalls (warps per issue): ' + ', '.join('%s=%.2f' % (n, x) for x, n in sorted(st, reverse=True)[:6]))
if ntop and len(rows) == 3:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]; si = h.index('Warp Stall Sampling (All Samples)'); s_ = h.index('Source')
    out = []; tot = 0
    for idx, r in enumerate(rows[2:]):
        try: x = float(r[si])
        except (ValueError, IndexError): continue
        tot += x; out.append((x, idx, r[s_].strip()[:80]))
    for x, idx, s in sorted(out, reverse=True)[:ntop]:
        print('  %6.2f%% of stall samples  sass#%4d %s' % (100 * x / max(tot, 1), idx, s))
