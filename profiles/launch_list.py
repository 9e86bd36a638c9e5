#!/usr/bin/env python
"""Per-kernel totals of one step from the ncu launch list (`--metrics gpu__time_duration.sum`, profiles/capture.sh pass 1).
Steps are delimited by the k_mcc launches (one per step): with `--steps 2 --warmup 2` the 3rd..4th k_mcc bracket a device-resident
step of the timed region and the 5th..6th an end-to-end step.  usage: python profiles/launch_list.py launches.csv "<bench command>" > rN_launches.md"""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
L = [(r[kn].split("(")[0].replace("void ", ""), float(r[mv].replace(",", "")) * scale[r[mu]]) for r in rows[1:] if len(r) > mv]
pos = [i for i, (n, _) in enumerate(L) if n == "k_mcc" or n.startswith("k_mcc<")]
print("# ncu launch list, B200, `%s`" % (sys.argv[2] if len(sys.argv) > 2 else "bench.py"))
print("# (`ncu --metrics gpu__time_duration.sum --clock-control none`; times are cold-cache and serialised: compare SHARES, not absolutes)")
print("# %d launches in the whole run; k_mcc launches (one per step) at positions %s" % (len(L), pos))
def block(title, a, b):
    seg = L[a:b]; tot = sum(t for _, t in seg); agg = {}
    for n, t in seg:
        e = agg.setdefault(n, [0.0, 0]); e[0] += t; e[1] += 1
    print("\n## %s: %d launches, %.2f ms of kernel time\n```" % (title, len(seg), tot / 1e3))
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%10.1f us %5.1f%%  x%-4d %s" % (t, 100 * t / tot, c, n))
    print("```")
if len(pos) >= 6:
    block("one device-resident step (timed region of `value`)", pos[2], pos[3])
    block("one end-to-end step (host buffers: inject, diagnostics, rho download)", pos[4], pos[5])
else:
    block("whole run", 0, len(L))
