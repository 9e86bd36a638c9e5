#!/usr/bin/env python
"""Kernel micro-benchmark: times the particle kernels one by one on a sorted synthetic species (CUDA-event timers of the
library).  usage: python profiles/microbench.py [mesh] [particles] [reps]"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
pg = importlib.import_module(bench.PKG + ".picgpu")
mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 128
npart = float(sys.argv[2]) if len(sys.argv) > 2 else 6e7
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
pg.init(0); pg.seed(1)
wl = bench.workload(mesh, npart * 2)          # species O gets half of the total
w = pg.World(mesh, mesh, mesh, wl["x0"], wl["xm"]); w.setTime(wl["dt"], 1 << 30)
for c, phi, sides in wl["rects"]:
    w.addRectangle(c, phi, sides)
w.computeObjectID()
sol = pg.PotentialSolver(w, 300, 1.0); sol.setReferenceValues(0, 0, 1e20); sol.solveGS(); sol.computeEF()
s = wl["species"][0]
neu = pg.Species("O", s["mass"], 0.0, w, s["mpw0"])
neu.reserve(int(s["count"] * 1.1)); neu.loadParticleBoxThermal(wl["box_c"], wl["box_s"], s["den"], s["T"]); neu.sort()
ele = pg.Species("e-", bench.ME, -bench.QE, w, s["mpw0"])
ele.reserve(int(s["count"] * 1.1)); ele.loadParticleBoxThermal(wl["box_c"], wl["box_s"], s["den"], 3000.0); ele.sort()
n = neu.getNumParticles(); ne = ele.getNumParticles()
neu.computeNumberDensity(); ele.computeNumberDensity()
neu.setDensityScale(neu.densityScale()); ele.setDensityScale(ele.densityScale())
dt = wl["dt"]
tests = {
    "push_heavy": (lambda: neu.advanceNonElectron(neu, neu, dt), 96, n),
    "push_heavy_deposit": (lambda: neu.advanceNonElectronDeposit(neu, neu, dt, count_cells=True), 104, n),
    "deposit_density": (lambda: neu.computeNumberDensity(), 32, n),
    "push_electrons": (lambda: ele.advanceElectrons(dt), 96, ne),
    "push_electrons_deposit": (lambda: ele.advanceElectronsDeposit(dt, count_cells=True), 104, ne),
    "count_per_cell": (lambda: (neu.advanceNonElectron(neu, neu, 0.0), neu.computeMacroParticlesCount()), 24, n),   # a push invalidates the cached count
}
only = os.environ.get("ONLY")
out = {}
for name, (fn, bytes_pp, cnt) in tests.items():
    if only and name not in only.split(","):
        continue
    fn(); pg.synchronize()
    pg.timers_reset(); pg.timers_enable(True)
    for _ in range(reps):
        fn()
    pg.synchronize(); pg.timers_enable(False)
    t = pg.timers_read()
    ms = t[name][0] / t[name][1]
    out[name] = dict(ms=round(ms, 4), GBps=round(bytes_pp * cnt / ms / 1e6, 1), frac=round(bytes_pp * cnt / ms / 1e6 / 6535.1, 3), n=cnt)
    print(name, out[name], flush=True)
print(json.dumps(out))
