"""Run-to-run reproducibility of the whole particle loop on the device.

The stochastic kernels draw from counter-based Philox streams addressed by (cell | particle | candidate, call), so their draws never
depend on scheduling.  What used to depend on it was the ORDER of the particles in the stores: products of MC collisions, particles
emitted by ions that neutralise on an electrode, addParticle's survivors and the hole filling of the compaction all took their slots
from atomic cursors - and the per-cell lists of the next MC call index those slots.  The stores are now ordered by construction
(mcc.cu: staged products in (cell, creation) order; push.cu: compaction plan from a slot bitmap; species.cu: candidates appended in
candidate order; step.cu: impacts in slot order, emitted particles placed by a prefix sum), so two runs from the same state and seed
must agree bit for bit, slot for slot - with every one of those paths exercised at once:

  * MC_MEX_Ionization with thousands of collisions and ionisations per step (split-off neutrals, new ions and electrons appended),
  * electrons absorbed by the electrodes and leaving through the open faces every step (compaction),
  * neutrals re-emitted diffusely from the electrodes, ions neutralised on the cathode (emission of neutrals into the neutral store),
  * a warm beam source and an explicit addParticles call with rejected candidates,
  * Species::merge, DSMC_MEX, the mover lists between re-sorts, a tail merge and a full re-sort.
"""
import numpy as np
import pytest

import util
from test_gpu_config4 import DT, E_ION, NI, NJ, NK, _common_phi, _initial_state

pytestmark = [pytest.mark.gpu]


def _run(picgpu, phi, neu, ele, ion, extra_e, seed, num_ts, merge_fraction):
    x0, xm, rects = util.discharge_geometry(NI, NJ, NK)
    w = util.build_world(picgpu.World, NI, NJ, NK, x0, xm, rects, dt=DT, num_ts=num_ts)
    mpw_n = float(neu[0, 6])
    O = picgpu.Species("O", 16 * util.AMU, 0.0, w, mpw_n, E_ION)
    Op = picgpu.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
    e = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    species = [O, Op, e]
    picgpu.seed(seed)
    picgpu.set_merge_fraction(merge_fraction)
    O.setParticles(neu); e.setParticles(ele); Op.setParticles(ion)
    tE, tS = util.momentum_transfer_table()
    mcc = picgpu.MC_MEX_Ionization(O, Op, e, w, tE, tS)
    mcc.setWsvMax(mpw_n * 8e-20 * 7e6)
    dsmc = picgpu.DSMC_MEX(O, w)
    src = picgpu.WarmBeamSource(e, w, 2e6, 5e15, 3000.0, "-x")
    sol = picgpu.PotentialSolver(w, 30, 1.0)
    sol.setReferenceValues(0.0, 0.0, 1e20)
    w.upload(picgpu.F_PHI, phi)
    sol.computeEF()
    rows = []
    for ts in range(1, num_ts + 1):
        src.sample()
        if ts == 2:
            e.addParticles(extra_e)                                       # half of these candidates lie inside an electrode or outside the box
        st = mcc.apply(DT)
        if ts % 2 == 0:
            dsmc.apply(DT)
        for sp in species:
            if sp is e:
                sp.advanceElectrons(DT)
            else:
                sp.advanceNonElectron(O, O, DT)
            sp.computeNumberDensity(); sp.computeMacroParticlesCount()
        if ts == 3:
            O.merge()
        w.computeChargeDensity(species)
        sol.solveGS(); sol.computeEF()
        rows.append((ts, st.candidates, st.collisions, st.ionizations, O.getNumParticles(), Op.getNumParticles(), e.getNumParticles(),
                     O.diagnostics()[2], e.diagnostics()[2], w.getPE()))
    final = {"O": O.getParticles(), "Op": Op.getParticles(), "e": e.getParticles(), "phi": w.phi, "rho": w.rho}
    stats = picgpu.mover_stats() if hasattr(picgpu, "mover_stats") else None
    for o in (src, dsmc, mcc, sol, O, Op, e, w):
        o.close()
    picgpu.set_merge_fraction(0.12)
    return rows, final, stats


@pytest.mark.parametrize("merge_fraction", [0.12, 0.01])
def test_two_runs_from_the_same_state_agree_slot_for_slot(picgpu, merge_fraction):
    num_ts = 7
    mpw_n = 5e12
    phi = _common_phi(picgpu)
    neu, ele = _initial_state(576_000, mpw_n, 64_000, ele_ev=(5.0, 120.0))
    x0, xm, _ = util.discharge_geometry(NI, NJ, NK)
    L = xm - x0
    rng = np.random.default_rng(7)
    # ions just above the cathode face (z = x0 + 0.05 Lz), falling onto it at 3e4 m/s: they cross 1-30 nm within the run and neutralise there
    n_ion = 20_000
    ion = np.empty((n_ion, 7))
    ion[:, 0:2] = x0[0:2] + rng.random((n_ion, 2)) * L[0:2] * 0.999
    ion[:, 2] = x0[2] + 0.05 * L[2] + rng.uniform(1e-12, 1.5e-7, n_ion)
    ion[:, 3:5] = rng.normal(0, 500.0, (n_ion, 2)); ion[:, 5] = -3e4
    ion[:, 6] = np.where(rng.random(n_ion) < 0.5, 100.0, 2.5 * mpw_n)    # the heavy ones emit 2-3 neutrals each (Species.cpp:225-232)
    extra_e = util.random_particles(30_000, x0 - 0.05 * L, xm + 0.05 * L, 3, vth=1e6, mpw=(100.0, 100.0))
    a_rows, a, _ = _run(picgpu, phi, neu, ele, ion, extra_e, 21, num_ts, merge_fraction)
    b_rows, b, _ = _run(picgpu, phi, neu, ele, ion, extra_e, 21, num_ts, merge_fraction)
    c_rows, c, _ = _run(picgpu, phi, neu, ele, ion, extra_e, 22, num_ts, merge_fraction)
    # every path did something
    assert a_rows[-1][2] > 1000 and sum(r[3] for r in a_rows) > 100                      # collisions, ionisations
    assert a_rows[-1][4] > 576_000 and a_rows[-1][5] < n_ion + sum(r[3] for r in a_rows)   # neutrals split off / emitted; ions absorbed on the cathode
    # identical, row by row and slot by slot (no sorting of the particle arrays)
    assert a_rows == b_rows
    for k in ("O", "Op", "e", "phi", "rho"):
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
    # and the comparison has teeth: another seed gives another history
    assert a_rows != c_rows and not (a["e"].shape == c["e"].shape and np.array_equal(a["e"], c["e"]))
