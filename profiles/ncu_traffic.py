#!/usr/bin/env python
"""DRAM traffic per launch of the hot kernels from `ncu --set full` reports -> profiles/ncu_traffic_<tag>.json (read by bench.py for
roofline.traffic).  usage: python profiles/ncu_traffic.py <tag> <report.ncu-rep> [...]"""
import csv, json, subprocess, sys
tag = sys.argv[1]
out = {}
for rep in sys.argv[2:]:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    kn = h.index('Kernel Name'); r_ = h.index('dram__bytes_read.sum'); w_ = h.index('dram__bytes_write.sum'); t_ = h.index('gpu__time_duration.sum')
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tscale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
    for v in rows[2:]:
        if len(v) <= max(r_, w_):
            continue
        name = v[kn].split('(')[0].replace('void ', '')
        b = float(v[r_]) * scale[u[r_]] + float(v[w_]) * scale[u[w_]]
        out.setdefault(name, []).append({"dram_bytes": b, "ms_under_ncu": float(v[t_]) * tscale[u[t_]]})
json.dump({"tag": tag, "how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/capture.sh)", "kernels": out},
          open('profiles/ncu_traffic_%s.json' % tag, 'w'), indent=1)
for k, v in out.items():
    print(k, [round(e["dram_bytes"] / 1e9, 3) for e in v])
