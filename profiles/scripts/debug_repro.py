"""Bisects run-to-run differences: the sequence of tests/test_gpu_reproducible.py, stores compared after every operation."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from test_gpu_config4 import DT, E_ION, NI, NJ, NK, _common_phi, _initial_state
pg = importlib.import_module("engineering-degree-in-plasma-simulations_b200.picgpu")
pg.init(0)

def run(phi, neu, ele, ion, extra_e, seed, num_ts, log):
    x0, xm, rects = util.discharge_geometry(NI, NJ, NK)
    w = util.build_world(pg.World, NI, NJ, NK, x0, xm, rects, dt=DT, num_ts=num_ts)
    mpw_n = float(neu[0, 6])
    O = pg.Species("O", 16 * util.AMU, 0.0, w, mpw_n, E_ION); Op = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0); e = pg.Species("e-", util.ME, -util.QE, w, 100.0)
    species = [O, Op, e]
    pg.seed(seed)
    O.setParticles(neu); e.setParticles(ele); Op.setParticles(ion)
    tE, tS = util.momentum_transfer_table()
    mcc = pg.MC_MEX_Ionization(O, Op, e, w, tE, tS); mcc.setWsvMax(mpw_n * 8e-20 * 7e6)
    dsmc = pg.DSMC_MEX(O, w)
    src = pg.WarmBeamSource(e, w, 2e6, 5e15, 3000.0, "-x")
    sol = pg.PotentialSolver(w, 30, 1.0); sol.setReferenceValues(0.0, 0.0, 1e20)
    w.upload(pg.F_PHI, phi); sol.computeEF()
    def snap(tag):
        log.append((tag, O.getParticles(), Op.getParticles(), e.getParticles(), w.phi.copy()))
    for ts in range(1, num_ts + 1):
        src.sample(); snap("ts%d source" % ts)
        if ts == 2:
            e.addParticles(extra_e); snap("ts%d add" % ts)
        mcc.apply(DT); snap("ts%d mcc" % ts)
        if ts % 2 == 0:
            dsmc.apply(DT); snap("ts%d dsmc" % ts)
        for sp, nm in zip(species, ("O", "Op", "e")):
            if sp is e: sp.advanceElectrons(DT)
            else: sp.advanceNonElectron(O, O, DT)
            snap("ts%d push %s" % (ts, nm))
            sp.computeNumberDensity(); sp.computeMacroParticlesCount()
        if ts == 3:
            O.merge(); snap("ts%d merge" % ts)
        w.computeChargeDensity(species); sol.solveGS(); sol.computeEF(); snap("ts%d fields" % ts)
    for o in (src, dsmc, mcc, sol, O, Op, e, w): o.close()

phi = _common_phi(pg)
mpw_n = 5e12
neu, ele = _initial_state(576_000, mpw_n, 64_000, ele_ev=(5.0, 120.0))
x0, xm, _ = util.discharge_geometry(NI, NJ, NK); L = xm - x0
rng = np.random.default_rng(7)
n_ion = 20_000
ion = np.empty((n_ion, 7))
ion[:, 0:2] = x0[0:2] + rng.random((n_ion, 2)) * L[0:2] * 0.999
ion[:, 2] = x0[2] + 0.05 * L[2] + rng.uniform(1e-12, 1.5e-7, n_ion)
ion[:, 3:5] = rng.normal(0, 500.0, (n_ion, 2)); ion[:, 5] = -3e4
ion[:, 6] = np.where(rng.random(n_ion) < 0.5, 100.0, 2.5 * mpw_n)
extra_e = util.random_particles(30_000, x0 - 0.05 * L, xm + 0.05 * L, 3, vth=1e6, mpw=(100.0, 100.0))
A, B = [], []
run(phi, neu, ele, ion, extra_e, 21, 3, A); run(phi, neu, ele, ion, extra_e, 21, 3, B)
for a, b in zip(A, B):
    res = []
    for k, nm in zip(range(1, 5), ("O", "Op", "e", "phi")):
        same = a[k].shape == b[k].shape and np.array_equal(a[k], b[k])
        if same: res.append(nm + ":same")
        elif a[k].shape == b[k].shape and nm != "phi" and np.array_equal(util.sort_rows(a[k]), util.sort_rows(b[k])): res.append(nm + ":ORDER(%d rows differ)" % int((a[k] != b[k]).any(axis=1).sum()))
        else: res.append(nm + ":VALUES %s %s" % (a[k].shape, b[k].shape))
    print(a[0].ljust(16), "  ".join(res))
