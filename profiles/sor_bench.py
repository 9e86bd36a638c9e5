"""SOR half-sweep micro-benchmark at 256^3 (scratch tool): python profiles/sor_bench.py [iterations]"""
import sys, importlib, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bench
pg = importlib.import_module("engineering-degree-in-plasma-simulations_b200.picgpu"); pg.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
wl = bench.workload(256, 1e9); m = wl["mesh"]
w = pg.World(m, m, m, wl["x0"], wl["xm"]); w.setTime(wl["dt"], 1 << 30)
for c, phi, sides in wl["rects"]:
    w.addRectangle(c, phi, sides)
w.computeObjectID()
sol = pg.PotentialSolver(w, 100, 1.0); sol.setReferenceValues(0.0, 0.0, 1e20)
sol.iterate(10)
pg.timers_reset(); pg.timers_enable(True)
sol.iterate(n)
pg.timers_enable(False)
t = pg.timers_read()
print({k: (round(v[0] / v[1] * 1e3, 2), v[1]) for k, v in t.items()}, "us per launch")
