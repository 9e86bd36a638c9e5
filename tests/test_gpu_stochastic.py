"""Stochastic device kernels (Philox streams) against the compiled reference (mt19937): ensemble statistics.

north_star: "Stochastic DSMC ionization must match the reference's ionization-rate and density moments within
stated ensemble confidence intervals."  Both sides run N_SEEDS independent seeds on identical inputs; for every
observable the two ensemble means must agree within CI_SIGMA combined standard errors (two-sample z test,
CI_SIGMA = 4.5 => false-alarm probability < 1e-5 per observable).
Deterministic sub-cases (no RNG involved) are compared exactly.
"""
import os
import tempfile

import numpy as np
import pytest

import util

pytestmark = [pytest.mark.gpu, pytest.mark.reference]

N_SEEDS = 32
CI_SIGMA = 4.5


def _agree(a, b, name, rel_floor=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    diff = abs(a.mean() - b.mean())
    assert diff <= CI_SIGMA * se + rel_floor * abs(b.mean()), f"{name}: gpu {a.mean():.6g} vs ref {b.mean():.6g}, diff {diff:.3g} > {CI_SIGMA} * {se:.3g}"


def _mcc_case(seed):
    ni, nj, nk = 7, 7, 9
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    rng = np.random.default_rng(1000 + seed)
    neu = util.random_particles(6000, x0, xm, seed=2000 + seed, vth=600.0, mpw=(5e11, 5e11), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    ele = util.random_particles(3000, x0, xm, seed=3000 + seed, vth=2.5e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    ele[:, 3:6] *= rng.uniform(0.5, 2.0, (len(ele), 1))      # 5..150 eV electrons: both elastic and ionising collisions
    return (ni, nj, nk, x0, xm, rects), neu, ele


def _energy_ev(parts):
    return 0.5 * util.ME * (parts[:, 3:6] ** 2).sum(1) / util.QE


def _run_mcc(mod, is_ref, seed, table, dt, wsv):
    (ni, nj, nk, x0, xm, rects), neu, ele = _mcc_case(0)          # identical inputs for every seed: only the RNG differs
    w = util.build_world(mod.World, ni, nj, nk, x0, xm, rects, dt=dt)
    E_ion = 1313.9 * 1000 / util.NA
    sn = mod.Species("O", 16 * util.AMU, 0.0, w, 5e11, E_ion)
    si = mod.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
    se = mod.Species("e-", util.ME, -util.QE, w, 100.0)
    sn.setParticles(neu); se.setParticles(ele)
    if is_ref:
        mod.seed(seed)
        m = mod.MC_MEX_Ionization(sn, si, se, w, table)
    else:
        mod.seed(seed)
        E, s = util.momentum_transfer_table()
        m = mod.MC_MEX_Ionization(sn, si, se, w, E, s)
    m.setWsvMax(wsv)
    m.apply(dt)
    pe, pi_, pn = se.getParticles(), si.getParticles(), sn.getParticles()
    out = dict(n_ion=len(pi_), n_split=len(pn) - len(neu), n_ele=len(pe) - len(ele),
               e_mean=_energy_ev(pe).mean(), e_new=_energy_ev(pe[len(ele):]).mean() if len(pe) > len(ele) else 0.0,
               w_neu=pn[:, 6].sum(), w_ion=pi_[:, 6].sum() if len(pi_) else 0.0, w_ele=pe[:, 6].sum(),
               ez_mean=pe[:, 5].mean(), ion_z=pi_[:, 2].mean() if len(pi_) else 0.0)
    for o in (m, sn, si, se, w):
        o.close()
    return out


def test_mc_ionization_ensemble(picgpu, ref):
    dt = 1e-10
    wsv = 5e11 * 8e-20 * 8e6          # a realistic ceiling W*sigma*g for these populations
    with tempfile.TemporaryDirectory() as d:
        table = util.write_table(os.path.join(d, "Oxygen_momentum_transfer.txt"))
        G = [_run_mcc(picgpu, False, s, table, dt, wsv) for s in range(N_SEEDS)]
        R = [_run_mcc(ref, True, 100 + s, table, dt, wsv) for s in range(N_SEEDS)]
    assert np.mean([r["n_ion"] for r in R]) > 20 and np.mean([r["n_split"] for r in R]) > 20     # the case exercises both branches
    for key in ("n_ion", "n_split", "n_ele", "e_mean", "e_new", "w_ion", "ez_mean", "ion_z"):
        _agree([g[key] for g in G], [r[key] for r in R], key)
    # exact invariants on every run: weight moved from neutrals to ions only by ionisation; electrons gain what ions gain
    for g in G:
        assert g["n_ion"] == g["n_ele"]
        assert abs(g["w_neu"] + g["w_ion"] - 6000 * 5e11) <= 1e-12 * 6000 * 5e11
        assert abs(g["w_ele"] - (3000 * 100.0 + g["w_ion"])) <= 1e-6


def test_mover_lists_are_exact_and_equivalent_to_resorting(picgpu, orc=None):
    """Between sorts the collision kernel works on a stale partition patched with mover lists.  (1) The per-cell list
    lengths must equal the per-cell particle counts exactly after pushes, deaths (hole filling) and appends.  (2) An MC
    ensemble run on patched lists must agree with the same ensemble run after a forced full sort."""
    pg = picgpu
    ni, nj, nk = 9, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    E, sg = util.momentum_transfer_table()
    E_ion = 1313.9 * 1000 / util.NA
    neu0 = util.random_particles(40000, x0, xm, seed=71, vth=600.0, mpw=(5e11, 5e11), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    ele0 = util.random_particles(20000, x0, xm, seed=72, vth=2.5e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    extra = util.random_particles(1500, x0, xm, seed=73, vth=2.5e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.2), hi_frac=(1, 1, 0.8))

    def run(seed, frac):
        pg.set_mover_fraction(frac); pg.seed(seed)
        w = util.build_world(pg.World, ni, nj, nk, x0, xm, rects, dt=1e-10)
        w.upload(pg.F_EF, util.smooth_ef((ni, nj, nk), x0, xm, seed=3, amp=2e6))
        sn = pg.Species("O", 16 * util.AMU, 0.0, w, 5e11, E_ion); si = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0); se = pg.Species("e-", util.ME, -util.QE, w, 100.0)
        sn.setParticles(neu0); se.setParticles(ele0)
        sn.sort(); se.sort()
        se.advanceElectrons(1.5e-11)                      # ~5 % of the electrons change cell, some are absorbed (holes filled from the tail)
        sn.advanceNonElectron(sn, sn, 2e-8)
        se.addParticles(extra)                            # appended beyond the partition
        m = pg.MC_MEX_Ionization(sn, si, se, w, E, sg)
        m.setWsvMax(5e11 * 8e-20 * 8e6)
        if frac > 0:
            se.computeMacroParticlesCount(); sn.computeMacroParticlesCount()
            assert np.array_equal(m.listCounts(1, w), se.macro_part_count)
            assert np.array_equal(m.listCounts(0, w), sn.macro_part_count)
        st = m.apply(1e-10)
        out = dict(coll=st.collisions, ion=st.ionizations, ne=se.getNumParticles(), nn=sn.getNumParticles(), ke=se.diagnostics()[2])
        for o in (m, sn, si, se, w):
            o.close()
        return out
    A = [run(s, 0.10) for s in range(N_SEEDS)]            # patched lists
    B = [run(100 + s, 0.0) for s in range(N_SEEDS)]       # full re-sort, as the reference does every step
    pg.set_mover_fraction(0.15)
    assert np.mean([b["coll"] for b in B]) > 100
    for key in ("coll", "ion", "ne", "nn", "ke"):
        _agree([a[key] for a in A], [b[key] for b in B], key)


def test_mover_lists_from_the_deposit_pass(picgpu):
    """Once per-cell lists are in use, the cell-partition deposit lists the movers as a by-product and the next list build
    only scans the appended tail.  The lists must stay exact through pushes, deaths, appends and MC products."""
    pg = picgpu
    ni, nj, nk = 9, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    E, sg = util.momentum_transfer_table()
    E_ion = 1313.9 * 1000 / util.NA
    neu0 = util.random_particles(40000, x0, xm, seed=81, vth=600.0, mpw=(5e11, 5e11), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    ele0 = util.random_particles(20000, x0, xm, seed=82, vth=2.5e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    pg.set_mover_fraction(0.4); pg.seed(5)
    w = util.build_world(pg.World, ni, nj, nk, x0, xm, rects, dt=1e-10)
    w.upload(pg.F_EF, util.smooth_ef((ni, nj, nk), x0, xm, seed=3, amp=2e6))
    sn = pg.Species("O", 16 * util.AMU, 0.0, w, 5e11, E_ion); si = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0); se = pg.Species("e-", util.ME, -util.QE, w, 100.0)
    sn.setParticles(neu0); se.setParticles(ele0)
    sn.sort(); se.sort()
    m = pg.MC_MEX_Ionization(sn, si, se, w, E, sg)
    m.setWsvMax(5e11 * 8e-20 * 8e6)
    m.apply(1e-10)                                        # first use of the lists: from now on the deposit lists the movers
    f0, s0, r0 = pg.mover_stats()
    for it in range(4):
        se.advanceElectrons(2e-12); se.computeNumberDensity()
        sn.advanceNonElectron(sn, sn, 6e-9); sn.computeNumberDensity()
        if it == 2:
            se.addParticles(util.random_particles(700, x0, xm, seed=90 + it, vth=2.5e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.2), hi_frac=(1, 1, 0.8)))   # invalidates the by-product
        se.computeMacroParticlesCount(); sn.computeMacroParticlesCount()
        assert np.array_equal(m.listCounts(1, w), se.macro_part_count)
        assert np.array_equal(m.listCounts(0, w), sn.macro_part_count)
        st = m.apply(1e-10)                               # appends ions / electrons (and split neutrals) beyond the partitions
        assert st.collisions > 0
    f1, s1, r1 = pg.mover_stats()
    # the deposit's list was re-used (MC products pile up beyond the partition, so a periodic full sort may also occur: r1 >= r0)
    assert f1 - f0 >= 3 and r1 >= r0
    pg.set_mover_fraction(0.15)
    for o in (m, sn, si, se, w):
        o.close()


def test_repeated_collision_calls_without_a_push_keep_the_lists_exact(picgpu):
    """Several MC calls in a row on species that are not pushed in between (the reference's subcycled loop: neutrals are advanced every
    100th step, the interaction runs every step): every list build starts from the deposit's mover list again and appends only the tail
    that the previous call created - no slot may be listed twice (sort.cu)."""
    pg = picgpu
    ni, nj, nk = 9, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    E, sg = util.momentum_transfer_table()
    E_ion = 1313.9 * 1000 / util.NA
    neu0 = util.random_particles(40000, x0, xm, seed=181, vth=600.0, mpw=(5e11, 5e11), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    ele0 = util.random_particles(20000, x0, xm, seed=182, vth=2.5e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    pg.set_mover_fraction(0.4); pg.set_merge_fraction(0.5); pg.seed(6)
    w = util.build_world(pg.World, ni, nj, nk, x0, xm, rects, dt=1e-10)
    w.upload(pg.F_EF, util.smooth_ef((ni, nj, nk), x0, xm, seed=3, amp=2e6))
    sn = pg.Species("O", 16 * util.AMU, 0.0, w, 5e11, E_ion); si = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0); se = pg.Species("e-", util.ME, -util.QE, w, 100.0)
    sn.setParticles(neu0); se.setParticles(ele0)
    sn.sort(); se.sort()
    m = pg.MC_MEX_Ionization(sn, si, se, w, E, sg)
    m.setWsvMax(5e11 * 8e-20 * 8e6)
    m.apply(1e-10)                                        # first use of the lists
    se.advanceElectrons(2e-12); se.computeNumberDensity()     # the deposit passes list the movers of both partitions
    sn.advanceNonElectron(sn, sn, 6e-9); sn.computeNumberDensity()
    n_neu = [sn.getNumParticles()]
    for it in range(4):                                   # no push, no deposit in between: the tails grow call by call
        st = m.apply(1e-10)
        assert st.collisions > 0 and st.dropped == 0
        n_neu.append(sn.getNumParticles())
        se.computeMacroParticlesCount(); sn.computeMacroParticlesCount()
        le, ln = m.listCounts(1, w), m.listCounts(0, w)
        assert np.array_equal(le, se.macro_part_count) and le.sum() == se.getNumParticles()
        assert np.array_equal(ln, sn.macro_part_count) and ln.sum() == sn.getNumParticles()
    assert all(b > a for a, b in zip(n_neu, n_neu[1:]))   # every call split neutrals off
    pg.set_mover_fraction(0.15); pg.set_merge_fraction(0.12)
    for o in (m, sn, si, se, w):
        o.close()


def test_tail_merge_keeps_particles_lists_and_deposit_exact(picgpu, orc):
    """Appended particles form a tail behind the cell partition; above the merge fraction the tail is merged into the partition
    (sort.cu: merge_tail) instead of re-sorting the store.  After each merge: the same multiset of particles, a partition that covers
    the whole store, exact per-cell lists (with drifted movers, deaths and a deposit-produced mover list in play) and a bit-exact deposit."""
    pg = picgpu
    ni, nj, nk = 9, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    E, sg = util.momentum_transfer_table()
    g = util.build_grid(orc, ni, nj, nk, x0, xm, rects)
    neu0 = util.random_particles(40000, x0, xm, seed=171, vth=600.0, mpw=(5e11, 5e11), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    ele0 = util.random_particles(30000, x0, xm, seed=172, vth=2.5e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    pg.set_mover_fraction(0.4); pg.set_merge_fraction(0.03); pg.seed(6)
    w = util.build_world(pg.World, ni, nj, nk, x0, xm, rects, dt=1e-10)
    w.upload(pg.F_EF, util.smooth_ef((ni, nj, nk), x0, xm, seed=3, amp=2e6))
    sn = pg.Species("O", 16 * util.AMU, 0.0, w, 5e11, 1313.9 * 1000 / util.NA); si = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0); se = pg.Species("e-", util.ME, -util.QE, w, 100.0)
    sn.setParticles(neu0); se.setParticles(ele0)
    sn.sort(); se.sort()
    m = pg.MC_MEX_Ionization(sn, si, se, w, E, sg)
    m.setWsvMax(5e11 * 8e-20 * 8e6)
    m.listCounts(1, w)                                    # lists in use: deposit passes list the movers from now on
    merges0 = pg.tail_merge_count()
    for it in range(4):
        se.advanceElectrons(2e-12)                        # some electrons change cell, some are absorbed (holes filled from the end of the store)
        se.computeNumberDensity()                         # cell-group deposit over the partition: lists the movers on the fly
        add = util.random_particles(2500 + 500 * it, x0, xm, seed=190 + it, vth=2.5e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.2), hi_frac=(1, 1, 0.8))
        se.addParticles(add)                              # > 3 % of the store appended behind the partition
        before = util.sort_rows(se.getParticles())
        assert se.partitionSize() < se.getNumParticles()
        se.computeMacroParticlesCount()
        assert np.array_equal(m.listCounts(1, w), se.macro_part_count)       # builds the lists: merges the tail first
        assert pg.tail_merge_count() >= merges0 + it + 1                      # (the collision call below may also merge the split-off neutrals)
        assert se.partitionSize() == se.getNumParticles()
        after = se.getParticles()
        assert np.array_equal(util.sort_rows(after), before)                  # nothing lost, nothing duplicated, nothing altered
        se.computeNumberDensity()                                             # cell-group deposit over the merged partition
        assert np.array_equal(se.den_fixed, g.deposit_fixed(after, se.densityScale()))
        se.computeMacroParticlesCount()
        assert np.array_equal(se.macro_part_count.ravel(), g.count_per_cell(after).ravel())
        assert np.array_equal(m.listCounts(1, w), se.macro_part_count)
        st = m.apply(1e-10)                               # the collision kernel on the merged partition
        assert st.collisions > 0
    pg.set_mover_fraction(0.15); pg.set_merge_fraction(0.12)
    for o in (m, sn, si, se, w):
        o.close()


def test_cross_sections_match_reference(picgpu, ref):
    x0, xm, rects = util.discharge_geometry(7, 7, 9)
    E, s = util.momentum_transfer_table()
    with tempfile.TemporaryDirectory() as d:
        table = util.write_table(os.path.join(d, "t.txt"))
        wr = util.build_world(ref.World, 7, 7, 9, x0, xm, rects)
        wg = util.build_world(picgpu.World, 7, 7, 9, x0, xm, rects)
        E_ion = 1313.9 * 1000 / util.NA
        sr = [ref.Species("O", 16 * util.AMU, 0.0, wr, 5e11, E_ion), ref.Species("O+", 16 * util.AMU, util.QE, wr, 100.0), ref.Species("e-", util.ME, -util.QE, wr, 100.0)]
        sg = [picgpu.Species("O", 16 * util.AMU, 0.0, wg, 5e11, E_ion), picgpu.Species("O+", 16 * util.AMU, util.QE, wg, 100.0), picgpu.Species("e-", util.ME, -util.QE, wg, 100.0)]
        mr = ref.MC_MEX_Ionization(sr[0], sr[1], sr[2], wr, table)
        mg = picgpu.MC_MEX_Ionization(sg[0], sg[1], sg[2], wg, E, s)
    q = np.concatenate([np.logspace(-4, 7, 300), E, [13.6, 13.618, 13.62, 0.0]])
    sc, si = mg.sigma(q)
    assert np.array_equal(sc, [mr.sigmaColl(e) for e in q])                  # table lookup + lerp: same bits
    ref_si = np.array([mr.sigmaIon(e) for e in q])
    assert np.allclose(si, ref_si, rtol=1e-13, atol=0)                       # log/exp differ by ulps between libm and CUDA
    # constructor preconditions (Interactions.cpp:479-489) surface as argument errors
    light = picgpu.Species("O", 16 * util.AMU, 0.0, wg, 50.0, E_ion)
    with pytest.raises(picgpu.PicgError):
        picgpu.MC_MEX_Ionization(light, sg[1], sg[2], wg, E, s)
    noion = picgpu.Species("O", 16 * util.AMU, 0.0, wg, 5e11)
    with pytest.raises(picgpu.PicgError):
        picgpu.MC_MEX_Ionization(noion, sg[1], sg[2], wg, E, s)
    for o in [mr, mg, light, noion] + sr + sg + [wr, wg]:
        o.close()


def _heavy_worlds(picgpu, ref, spheres=()):
    ni, nj, nk = 11, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    wr = util.build_world(ref.World, ni, nj, nk, x0, xm, rects, spheres, dt=1e-9)
    wg = util.build_world(picgpu.World, ni, nj, nk, x0, xm, rects, spheres, dt=1e-9)
    ef = util.smooth_ef((ni, nj, nk), x0, xm, seed=2, amp=2e5)
    wr.set(3, ef); wg.upload(picgpu.F_EF, ef)
    return wr, wg, x0, xm


def test_heavy_push_ions_deterministic_part(picgpu, ref):
    """Ions: kick, drift, absorption on electrodes / sphere / walls involve no RNG (the injected-neutral count
    int(100/5e11 + rnd()) is 0), so the survivors must match the reference bit for bit."""
    sph = [((0.0, 0.0, 0.0025), -100.0, 0.0007)]
    wr, wg, x0, xm = _heavy_worlds(picgpu, ref, sph)
    parts = util.random_particles(40000, x0, xm, seed=5, vth=4e4, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.1), hi_frac=(1, 1, 0.9))
    out = []
    for mod, w in ((ref, wr), (picgpu, wg)):
        mod.seed(7)
        neu = mod.Species("O", 16 * util.AMU, 0.0, w, 5e11)
        ion = mod.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
        ion.setParticles(parts)
        for _ in range(3):
            ion.advanceNonElectron(neu, neu, 2e-8)
        out.append((ion.getParticles(), neu.getNumParticles()))
        ion.close(); neu.close()
    (pr, nr), (pg, ng) = out
    assert 0 < len(pr) < len(parts) and nr == 0 and ng == 0
    assert np.array_equal(util.sort_rows(pg), util.sort_rows(pr))
    wr.close(); wg.close()


def test_heavy_push_neutral_reflection_statistics(picgpu, ref):
    """Neutrals are re-emitted diffusely from surfaces (sampleReflectedVelocity, 11 RNG draws per bounce)."""
    sph = [((0.0, 0.0, 0.0025), -100.0, 0.0009)]
    wr, wg, x0, xm = _heavy_worlds(picgpu, ref, sph)
    parts = util.random_particles(8000, x0, xm, seed=9, vth=900.0, mpw=(5e11, 5e11), lo_frac=(0.1, 0.1, 0.12), hi_frac=(0.9, 0.9, 0.88))
    dt = 4e-7                                  # ~0.4 mm per step: a good fraction reaches a surface
    res = {"g": [], "r": []}
    for tag, mod, w in (("r", ref, wr), ("g", picgpu, wg)):
        for s in range(N_SEEDS):
            mod.seed(50 + s + (0 if tag == "g" else 500))
            neu = mod.Species("O", 16 * util.AMU, 0.0, w, 5e11)
            neu.setParticles(parts)
            neu.advanceNonElectron(neu, neu, dt)
            p = neu.getParticles()
            moved = ~np.isin(p[:, 3], parts[:, 3])        # velocity changed <=> the particle bounced
            res[tag].append(dict(n=len(p), n_bounced=int(moved.sum()), speed=np.linalg.norm(p[moved, 3:6], axis=1).mean(),
                                 ke=(p[:, 3:6] ** 2).sum(), z=p[:, 2].mean(), vz_b=p[moved, 5].mean()))
            neu.close()
    assert np.mean([r["n_bounced"] for r in res["r"]]) > 100
    for key in ("n", "n_bounced", "speed", "ke", "z", "vz_b"):
        _agree([g[key] for g in res["g"]], [r[key] for r in res["r"]], key)
    wr.close(); wg.close()


def test_sources_statistics(picgpu, ref):
    ni, nj, nk = 11, 9, 13
    x0, xm = np.array([-0.1, -0.1, 0.0]), np.array([0.1, 0.1, 0.4])
    for face, T in (("-z", None), ("+x", 1000.0), ("y+", None)):
        res = {"g": [], "r": []}
        for tag, mod in (("r", ref), ("g", picgpu)):
            w = util.build_world(mod.World, ni, nj, nk, x0, xm, dt=1e-7)
            sp = mod.Species("O+", 16 * util.AMU, util.QE, w, 2e2)
            mod.seed(11 if tag == "g" else 12)
            if tag == "r":
                src = ref.Source(sp, w, 7000.0, 1e10, face, T)
            else:
                src = picgpu.WarmBeamSource(sp, w, 7000.0, 1e10, T, face) if T else picgpu.ColdBeamSource(sp, w, 7000.0, 1e10, face)
            for _ in range(N_SEEDS):
                before = sp.getNumParticles()
                src.sample()
                p = sp.getParticles()[before:]
                res[tag].append(dict(n=len(p), x=p[:, 0].mean(), y=p[:, 1].mean(), z=p[:, 2].mean(), u=p[:, 3].mean(), v=p[:, 4].mean(), w=p[:, 5].mean(),
                                     ke=(p[:, 3:6] ** 2).sum(1).mean(), mpw=p[:, 6].mean()))
            src.close(); sp.close(); w.close()
        assert res["r"][0]["n"] > 500
        for key in res["r"][0]:
            _agree([g[key] for g in res["g"]], [r[key] for r in res["r"]], f"{face}:{key}", rel_floor=1e-12)


def test_thermal_loader_statistics(picgpu, ref):
    ni, nj, nk = 11, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    xc = 0.5 * (x0 + xm); L = xm - x0
    out = {}
    for tag, mod in (("r", ref), ("g", picgpu)):
        w = util.build_world(mod.World, ni, nj, nk, x0, xm, rects)
        sp = mod.Species("e-", util.ME, -util.QE, w, 1e5)
        mod.seed(3)
        sp.loadParticleBoxThermal(xc, (L[0], L[1], L[2] * 0.9), 1e16, 3000.0)
        out[tag] = sp.getParticles()
        sp.close(); w.close()
    g, r = out["g"], out["r"]
    assert abs(len(g) - len(r)) < 5 * np.sqrt(len(r) * 0.2) + 5          # same request, same in-electrode rejection fraction
    for c in range(6):                                                    # position and velocity moments
        for f, nm in ((np.mean, "mean"), (np.std, "std")):
            a, b = f(g[:, c]), f(r[:, c])
            scale = np.std(r[:, c])
            assert abs(a - b) < 6 * scale / np.sqrt(len(r)), (c, nm, a, b)
    sg, sr = np.linalg.norm(g[:, 3:6], axis=1), np.linalg.norm(r[:, 3:6], axis=1)
    assert abs(sg.mean() - sr.mean()) < 6 * sr.std() / np.sqrt(len(sr))
    assert np.all(g[:, 6] == 1e5)


def test_mc_candidates_are_dealt_out_over_ranks(picgpu):
    """Multi-GPU rule, emulated on one device: G ranks each hold every G-th particle.  A cell's candidate count is
    estimated from the local populations (x G^2), rounded once like the reference's (Interactions.cpp:646-647) and dealt out
    to the ranks.  In a regime of about one candidate per cell and step the ranks together must try as many pairs as a single
    rank holding everything (rounding per rank would try almost none)."""
    pg = picgpu
    (ni, nj, nk, x0, xm, rects), neu, ele = _mcc_case(0)
    E, sg = util.momentum_transfer_table()
    E_ion = 1313.9 * 1000 / util.NA
    dt, wsv = 2e-11, 5e11 * 8e-20 * 8e6

    def run(rank, G):
        pg.set_rank(rank, G); pg.seed(77)
        w = util.build_world(pg.World, ni, nj, nk, x0, xm, rects, dt=dt)
        sn = pg.Species("O", 16 * util.AMU, 0.0, w, 5e11, E_ion); si = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0); se = pg.Species("e-", util.ME, -util.QE, w, 100.0)
        sn.setParticles(neu[rank::G]); se.setParticles(ele[rank::G])
        m = pg.MC_MEX_Ionization(sn, si, se, w, E, sg); m.setWsvMax(wsv)
        st = m.apply(dt)
        out = (st.candidates, st.collisions)
        for o in (m, sn, si, se, w):
            o.close()
        return out

    try:
        one = run(0, 1)
        G = 4
        parts = [run(r, G) for r in range(G)]
    finally:
        pg.set_rank(0, 1)
    cells = (ni - 1) * (nj - 1) * (nk - 1)
    assert 0.5 * cells < one[0] < 3 * cells                          # the regime the rule matters in: ~1 candidate per cell
    tot = sum(p[0] for p in parts)
    assert 0.85 * one[0] < tot < 1.15 * one[0], (one, parts)
    assert max(p[0] for p in parts) < 0.4 * one[0]                   # and they are spread over the ranks


def test_mc_ionization_one_cell_pair_by_pair(picgpu, orc):
    """Deterministic check below the ensemble level.  One cell; the device's Philox stream is replayed on the host and drives
    tests/mcc_restatement.py - the reference's candidate loop with collide / newVelocityElecton / cross-sections that are pinned bit
    for bit against the compiled reference on CPU (test_oracle_vs_reference.py).  Same candidates, same collisions, same
    ionisations, same products in the same order; values agree to the last digits (device libm vs host libm)."""
    import mcc_restatement as R
    from test_gpu_dsmc import _philox_stream
    pg = picgpu
    x0, xm = np.array([0.0, 0.0, 0.0]), np.array([2e-3, 2e-3, 2e-3])            # 3 nodes per axis: 8 cells of 1 mm, everything in cell 0
    seed, dt, wsv = 987, 7e-11, 5e11 * 8e-20 * 8e6
    E_ion = 1313.9 * 1000 / util.NA
    rng = np.random.default_rng(12)
    neu = util.random_particles(60, x0, 0.5 * xm, seed=31, vth=600.0, mpw=(5e11, 5e11)); neu[:, 6] = 5e11
    ele = util.random_particles(30, x0, 0.5 * xm, seed=32, vth=1.0, mpw=(100.0, 100.0)); ele[:, 6] = 100.0
    ele[:, 3:6] *= (np.sqrt(2 * rng.uniform(5.0, 150.0, 30) * util.QE / util.ME) / np.linalg.norm(ele[:, 3:6], axis=1))[:, None]
    w = util.build_world(pg.World, 3, 3, 3, x0, xm, dt=dt)
    sn = pg.Species("O", 16 * util.AMU, 0.0, w, 5e11, E_ion); si = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0); se = pg.Species("e-", util.ME, -util.QE, w, 100.0)
    sn.setParticles(neu); se.setParticles(ele)
    pg.seed(seed)
    tE, tS = util.momentum_transfer_table()
    m = pg.MC_MEX_Ionization(sn, si, se, w, tE, tS); m.setWsvMax(wsv)
    st = m.apply(dt)
    got_n, got_e, got_i = sn.getParticles(), se.getParticles(), si.getParticles()

    M = R.MccModel(16 * util.AMU, util.ME, E_ion, tE, tS, 1e-3 * 1e-3 * 1e-3, 5e11, 100.0)
    ln, le = [list(map(float, r)) for r in neu], [list(map(float, r)) for r in ele]
    stream = _philox_stream(orc, seed, 4 + 16 * 0, 0, 1)                          # RNG_MCC = 4, neutral species 0, rank 0; cell 0, first call
    cand, coll, ions, new_e, split, step_max = M.apply_cell(stream, ln, le, dt, wsv)
    assert cand > 30 and coll > 5 and len(split) > 0                              # the case exercises the loop
    assert (st.candidates, st.collisions, st.ionizations) == (cand, coll, len(ions))
    want_n, want_e = np.array(ln + split), np.array(le + new_e)
    assert got_n.shape == want_n.shape and got_e.shape == want_e.shape and len(got_i) == len(ions)
    assert np.allclose(got_n, want_n, rtol=1e-11, atol=0, equal_nan=True)
    assert np.allclose(got_e, want_e, rtol=1e-11, atol=1e-6, equal_nan=True)
    if ions:
        assert np.allclose(got_i, np.array(ions), rtol=1e-11, atol=0, equal_nan=True)
    assert abs(st.w_sigma_v_max - step_max) <= 1e-12 * step_max                   # the ceiling of the next call (:751-756)
    for o in (m, sn, si, se, w):
        o.close()


def test_loader_and_sources_replayed_particle_by_particle(picgpu, orc):
    """The thermal loader and the beam sources with the device's Philox streams replayed on the host through
    tests/sampler_restatement.py (sampleV3th pinned bit for bit against the compiled reference on CPU): the same candidates,
    the same addParticle filter, the same particles (as multisets: the append order depends on atomics)."""
    import sampler_restatement as S
    from test_gpu_dsmc import _philox_stream
    pg = picgpu
    seed = 2718
    # --- Species::loadParticleBoxThermal (Species.cpp:560-598) between the electrodes of the discharge geometry
    ni, nj, nk = 11, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    w = util.build_world(pg.World, ni, nj, nk, x0, xm, rects)
    g = util.build_grid(orc, ni, nj, nk, x0, xm, rects)
    pg.seed(seed)
    sp = pg.Species("e-", util.ME, -util.QE, w, 1e5)
    xc, L = 0.5 * (x0 + xm), xm - x0
    sides = np.array([L[0], L[1], L[2] * 0.96])                             # reaches into both electrodes (each 0.05 Lz deep): some candidates are rejected
    T, den = 3000.0, 6e15
    sp.loadParticleBoxThermal(xc, sides, den, T)
    got = util.sort_rows(sp.getParticles())
    n = int(den * (sides[0] * sides[1] * sides[2]) / 1e5)
    lo, hi = xc - sides / 2, xc + sides / 2
    want = []
    for t in range(n):
        r = _philox_stream(orc, seed, 1 + 16 * 0, t, 1)                     # RNG_LOADER = 1, species 0; particle index, first load call
        pos = [float(lo[a] + next(r) * (hi[a] - lo[a])) for a in range(3)]
        vel = S.sample_v3th(r, T, util.ME)
        if g.in_bounds(pos) and not g.in_object(pos):                       # Species::addParticle :420-434 (E == 0: the rewind is a no-op)
            want.append(pos + vel + [1e5])
    want = util.sort_rows(np.array(want))
    assert 1000 < len(want) < n and got.shape == want.shape
    assert np.array_equal(got[:, [0, 1, 2, 6]], want[:, [0, 1, 2, 6]])
    assert np.allclose(got[:, 3:6], want[:, 3:6], rtol=1e-12, atol=1e-6)
    sp.close(); w.close()
    # --- WarmBeamSource / ColdBeamSource::sample (Source.cpp:38-103,111-191) on a "-" and a "+" face
    x0, xm = np.array([-0.1, -0.1, 0.0]), np.array([0.1, 0.1, 0.4])
    dxs = (xm - x0) / (np.array([ni, nj, nk]) - 1)
    for face, axis, plus, Tsrc in (("-z", 2, False, 1000.0), ("+x", 0, True, None)):
        w = util.build_world(pg.World, ni, nj, nk, x0, xm, dt=1e-7)
        pg.seed(seed)
        sp = pg.Species("O+", 16 * util.AMU, util.QE, w, 2e2)
        src = pg.WarmBeamSource(sp, w, 7000.0, 1e10, Tsrc, face) if Tsrc else pg.ColdBeamSource(sp, w, 7000.0, 1e10, face)
        n_inj = src.sample()
        got = util.sort_rows(sp.getParticles())
        Ls = xm - x0
        A = Ls[1] * Ls[2] if axis == 0 else (Ls[0] * Ls[2] if axis == 1 else Ls[0] * Ls[1])
        num = 1e10 * 7000.0 * A * 1e-7 / 2e2
        r0 = _philox_stream(orc, seed, 2 + 16 * 0, 0xFFFFFFFFFFFFFFFF, 1)   # RNG_SOURCE = 2; the count draw uses index ~0
        n_macro = int(num + next(r0))                                       # Source.cpp:41
        if plus:
            Ls = Ls.copy(); Ls[axis] -= dxs[axis]                           # :51,69,86
        want = []
        for t in range(n_macro):
            r = _philox_stream(orc, seed, 2 + 16 * 0, t, 1)
            v = S.sample_v3th(r, Tsrc, 16 * util.AMU) if Tsrc else [0.0, 0.0, 0.0]
            v[axis] += -7000.0 if plus else 7000.0
            p = [0.0, 0.0, 0.0]
            for a in range(3):
                p[a] = float(x0[a] + Ls[a]) if (a == axis and plus) else (float(x0[a]) if a == axis else float(x0[a] + next(r) * Ls[a]))
            if all(x0[a] <= p[a] < xm[a] for a in range(3)):
                want.append(p + v + [2e2])
        want = util.sort_rows(np.array(want))
        assert n_inj == len(want) > 500 and got.shape == want.shape, (face, n_inj, len(want))
        assert np.array_equal(got[:, [0, 1, 2, 6]], want[:, [0, 1, 2, 6]])
        assert np.allclose(got[:, 3:6], want[:, 3:6], rtol=1e-12, atol=1e-6)
        src.close(); sp.close(); w.close()


def test_neutral_wall_reflection_replayed_particle_by_particle(picgpu, orc):
    """Diffuse re-emission of neutrals from the electrodes (Species.cpp:194-249, sampleReflectedVelocity :835-853) with the device's
    Philox stream (one per particle slot) replayed on the host through tests/heavy_restatement.py, which is pinned bit for bit against
    the compiled reference on CPU: the same particles survive, bounce (several times where the reference does) and leave."""
    import heavy_restatement as H
    from test_gpu_dsmc import _philox_stream
    from test_oracle_vs_reference import _heavy_case
    pg = picgpu
    x0, xm, rects, boxes = _heavy_case()
    w = util.build_world(pg.World, 11, 9, 13, x0, xm, rects)
    g = util.build_grid(orc, 11, 9, 13, x0, xm, rects)
    mass, dt, seed, n = 16 * util.AMU, 4e-7, 1618, 3000
    rng = np.random.default_rng(9)
    p = np.zeros((n, 7))
    p[:, 0:2] = x0[0:2] + rng.random((n, 2)) * (xm[0:2] - x0[0:2])
    low = np.arange(n) % 2 == 0
    p[:, 2] = np.where(low, 0.05 * 0.005 + rng.random(n) * 4e-4, 0.95 * 0.005 - rng.random(n) * 4e-4)      # just outside an electrode
    p[:, 3:6] = rng.normal(0, 700.0, (n, 3))
    p[:, 5] = np.where(low, -np.abs(p[:, 5]) - 300.0, np.abs(p[:, 5]) + 300.0)                              # heading into it
    p[:, 6] = 5e11
    keep = np.array([not g.in_object(r[0:3]) and bool(g.in_bounds(r[0:3])) for r in p])
    p = p[keep]
    pg.seed(seed)
    sp = pg.Species("O", mass, 0.0, w, 5e11)
    sp.setParticles(p)
    sp.advanceNonElectron(sp, sp, dt)
    got = util.sort_rows(sp.getParticles())
    want, n_bounced = [], 0
    for slot, r in enumerate(p):
        res = H.advance_neutral(_philox_stream(orc, seed, 3 + 16 * 0, slot, 1), g, boxes, mass, r[0:3], r[3:6], dt)   # RNG_HEAVY = 3, species 0, first heavy push
        if res is not None:
            want.append(list(res[0]) + list(res[1]) + [5e11])
            n_bounced += int(list(res[1]) != [float(c) for c in r[3:6]])
    want = util.sort_rows(np.array(want))
    assert n_bounced > 1000 and len(want) < len(p)                              # most hit an electrode, some left through the sides
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-11, atol=1e-15)
    sp.close(); w.close()


def test_ion_neutralisation_replayed_particle_by_particle(picgpu, orc):
    """Ions neutralised on the electrodes (Species.cpp:225-232): int(mpw / neutrals.mpw0 + rnd()) neutrals re-emitted from the hit point
    through addParticle.  Device Philox streams replayed through heavy_restatement.advance_ion (pinned on CPU): the same ions are
    absorbed, the same neutrals appear - including which of them addParticle rejects on the surface."""
    import heavy_restatement as H
    from test_gpu_dsmc import _philox_stream
    from test_oracle_vs_reference import _heavy_case
    pg = picgpu
    x0, xm, rects, boxes = _heavy_case()
    w = util.build_world(pg.World, 11, 9, 13, x0, xm, rects)
    g = util.build_grid(orc, 11, 9, 13, x0, xm, rects)
    mass, dt, seed, n = 16 * util.AMU, 4e-7, 577, 3000
    rng = np.random.default_rng(19)
    p = np.zeros((n, 7))
    p[:, 0:2] = x0[0:2] + rng.random((n, 2)) * (xm[0:2] - x0[0:2])
    low = np.arange(n) % 2 == 0
    p[:, 2] = np.where(low, 0.05 * 0.005 + rng.random(n) * 4e-4, 0.95 * 0.005 - rng.random(n) * 4e-4)
    p[:, 3:6] = rng.normal(0, 700.0, (n, 3))
    p[:, 5] = np.where(low, -np.abs(p[:, 5]) - 300.0, np.abs(p[:, 5]) + 300.0)
    p[:, 6] = 260.0
    keep = np.array([not g.in_object(r[0:3]) and bool(g.in_bounds(r[0:3])) for r in p])
    p = p[keep]
    pg.seed(seed)
    neu = pg.Species("O", mass, 0.0, w, 100.0)                                     # species 0
    ion = pg.Species("O+", mass, util.QE, w, 260.0)                                # species 1: its stream is RNG_HEAVY + 16
    ion.setParticles(p)
    ion.advanceNonElectron(neu, neu, dt)                                           # E == 0: no kick, no rewind
    got_i, got_n = util.sort_rows(ion.getParticles()), util.sort_rows(neu.getParticles())
    want_i, want_n = [], []
    for slot, r in enumerate(p):
        left, emitted = H.advance_ion(_philox_stream(orc, seed, 3 + 16 * 1, slot, 1), g, boxes, mass, 260.0, 100.0, r[0:3], r[3:6], dt)
        if left is not None:
            want_i.append(list(left[0]) + list(left[1]) + [260.0])
        want_n += emitted
    want_i, want_n = util.sort_rows(np.array(want_i)), util.sort_rows(np.array(want_n))
    assert len(want_n) > 1000 and len(want_i) < len(p) // 2
    assert got_i.shape == want_i.shape and got_n.shape == want_n.shape
    assert np.array_equal(got_i, want_i)                                           # free flight: no transcendental involved
    assert np.array_equal(got_n[:, [0, 1, 2, 6]], want_n[:, [0, 1, 2, 6]])
    assert np.allclose(got_n[:, 3:6], want_n[:, 3:6], rtol=1e-11, atol=1e-12)
    ion.close(); neu.close(); w.close()
