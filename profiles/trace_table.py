#!/usr/bin/env python
"""Per-step kernel-time table from a PICG_STEP_TRACE=1 run of bench.py (stderr).  usage: python profiles/trace_table.py bench.err"""
import json, sys
import numpy as np
rows = []
for line in open(sys.argv[1]):
    if line.startswith('step '):
        ts = int(line.split()[1].rstrip(':'))
        rows.append((ts, json.loads(line[line.index('ms=') + 3:])))
keys = ['mc_ionize', 'deposit_density', 'deposit_tail', 'sort_keys', 'sort_hist', 'sort_scan', 'sort_scatter', 'sort_permute', 'cell_start', 'push_neutral', 'push_electrons', 'push_heavy', 'sor_redblack']
SORT = ['sort_keys', 'sort_hist', 'sort_scan', 'sort_scatter', 'sort_permute', 'cell_start']
print('ts   ' + ' '.join('%9s' % k[:9] for k in keys) + '   sortsum   total')
for ts, ms in rows:
    if not ms.get('push_electrons'):
        continue
    print('%3d  ' % ts + ' '.join('%9.2f' % ms.get(k, 0) for k in keys) + '   %6.2f  %6.2f' % (sum(ms.get(k, 0) for k in SORT), sum(v for k, v in ms.items() if k != 'diagnostics')))
sel = [ms for ts, ms in rows if 6 <= ts <= 25]
print('steps 6-25: sort+lists %.2f ms/step, deposit_tail %.2f, mc_ionize %.2f, deposit_density %.2f, all kernels %.2f' % (
    np.mean([sum(ms.get(k, 0) for k in SORT) for ms in sel]), np.mean([ms.get('deposit_tail', 0) for ms in sel]), np.mean([ms.get('mc_ionize', 0) for ms in sel]),
    np.mean([ms.get('deposit_density', 0) for ms in sel]), np.mean([sum(v for k, v in ms.items() if k != 'diagnostics') for ms in sel])))
