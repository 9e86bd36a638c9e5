import sys, importlib; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, util
pg = importlib.import_module("engineering-degree-in-plasma-simulations_b200.picgpu"); pg.init(0)
ni, nj, nk = 11, 9, 13
x0, xm, rects = util.discharge_geometry(ni, nj, nk)
wg = util.build_world(pg.World, ni, nj, nk, x0, xm, rects)
ef = util.smooth_ef((ni, nj, nk), x0, xm, seed=5, amp=3e6)
wg.upload(pg.F_EF, ef)
parts = util.random_particles(50000, x0, xm, seed=43, vth=3e3, mpw=(5e11, 5e11), lo_frac=(0, 0, 0.2), hi_frac=(1, 1, 0.8))
ng = pg.Species("O", 16 * util.AMU, 0.0, wg, 5e11); ng.setParticles(parts)
ng.advanceNonElectron(ng, ng, 4e-8)
print("n after 1:", ng.getNumParticles())
ng.advanceNonElectron(ng, ng, 4e-8)
got = ng.getParticles()
p = parts.copy(); p[:,0:3] += p[:,3:6]*4e-8
alive = ((p[:,0:3] >= x0) & (p[:,0:3] < xm)).all(1)
p = p[alive]; print("numpy after 1:", len(p)); p[:,0:3] += p[:,3:6]*4e-8
alive = ((p[:,0:3] >= x0) & (p[:,0:3] < xm)).all(1)
print(len(got), alive.sum())
a=util.sort_rows(got); b=util.sort_rows(p[alive])
if a.shape==b.shape:
    print(np.array_equal(a,b), [(a[:,c]!=b[:,c]).sum() for c in range(7)])
    bad=np.where((a!=b).any(1))[0][:3]; print(a[bad]); print(b[bad])
