"""BASELINE config 3: the fixed-weight MC_MEX_Ionization of ch4/v2 (ch4/v2/Interactions.cpp:476-735), variant 1 of the device's
collision kernel (csrc/mcc.cu, picg_mcc_set_variant).

Three levels, as for the v3 algorithm: (1) tests/mcc_restatement.py::MccModelV2 pinned bit for bit against the compiled ch4/v2
reference on CPU (cross-sections, collide with the reference's own draws, a whole apply() on a one-cell case); (2) the device kernel
against that restatement pair by pair, its Philox stream replayed on the host; (3) 32-seed ensembles, device against compiled
reference (two-sample z test, 4.5 sigma)."""
import os
import tempfile

import numpy as np
import pytest

import util

E_ION = 1313.9 * 1000 / util.NA
N_SEEDS = 32
CI_SIGMA = 4.5


def _model(R, dv, mpw=(5e11, 100.0, 100.0)):
    tE, tS = util.momentum_transfer_table()
    return R.MccModelV2(16 * util.AMU, util.ME, E_ION, tE, tS, dv, *mpw)


def _species(mod, w, mpw=(5e11, 100.0, 100.0)):
    return (mod.Species("O", 16 * util.AMU, 0.0, w, mpw[0], E_ION), mod.Species("O+", 16 * util.AMU, util.QE, w, mpw[1]),
            mod.Species("e-", util.ME, -util.QE, w, mpw[2]))


def _electrons_ev(n, x0, xm, seed, lo_ev, hi_ev):
    rng = np.random.default_rng(seed)
    ele = util.random_particles(n, x0, xm, seed=seed + 1, vth=1.0, mpw=(100.0, 100.0)); ele[:, 6] = 100.0
    ele[:, 3:6] *= (np.sqrt(2 * rng.uniform(lo_ev, hi_ev, n) * util.QE / util.ME) / np.linalg.norm(ele[:, 3:6], axis=1))[:, None]
    return ele


# ----------------------------------------------------------------------------------------------- (1) restatement pinned on CPU
@pytest.mark.reference
def test_v2_restatement_pinned_against_reference(ref2, orc, tmp_path):
    import mcc_restatement as R
    ni, nj, nk = 3, 3, 3
    x0, xm = np.array([0.0, 0.0, 0.0]), np.array([2e-3, 2e-3, 2e-3])
    dt = 1.8e-10
    w = util.build_world(ref2.World, ni, nj, nk, x0, xm, dt=dt)
    ef = util.smooth_ef((ni, nj, nk), x0, xm, seed=5, amp=3e6)
    w.setEF(ef)
    sn, si, se = _species(ref2, w)
    table = util.write_table(str(tmp_path / "Oxygen_momentum_transfer.txt"))
    m = ref2.MC_MEX_Ionization(sn, si, se, w, table)
    M = _model(R, 1e-9)
    # cross-sections: the table lookup and the ionisation fit WITHOUT the threshold guard of v3 (negative below 9.793 eV)
    for E in np.exp(np.random.default_rng(1).uniform(np.log(1e-4), np.log(1e7), 400)):
        assert M.sigma_coll(float(E)) == m.sigmaColl(float(E))
        assert M.sigma_ion(float(E)) == m.sigmaIon(float(E))
    assert m.sigmaIon(5.0) < 0 and m.sigmaIon(12.0) > 0 and M.w_max0 == m.getWsvMax() == 1e-14
    # collide with the reference's own draws, sub-threshold ionisations (NaN ejected electron) included
    rng = np.random.default_rng(2)
    n_ion = n_nan = 0
    for k in range(400):
        vn = rng.normal(0, 600.0, 3)
        ev = rng.uniform(9.9, 16.0) if k % 4 == 0 else rng.uniform(10.0, 200.0)
        ve = rng.normal(0, 1.0, 3); ve *= np.sqrt(2 * ev * util.QE / util.ME) / np.linalg.norm(ve)
        s_coll = M.sigma_coll(M.E_rel_eV * float(np.sum((vn - ve) ** 2)))
        if k % 4 == 0:
            s_coll *= 1e-3                                            # force the ionising branch, also below the threshold (9.793 .. 13.618 eV)
        ref2.seed(5000 + k); draws = [ref2.rnd() for _ in range(8)]; ref2.seed(5000 + k)
        ion_r, vn_r, ve_r, vnew_r = m.collide(vn, ve, s_coll)
        ion_p, ve_p, vnew_p = M.collide(iter(draws), [float(x) for x in vn], [float(x) for x in ve], s_coll)
        assert ion_p == ion_r and np.array_equal(vn_r, vn)
        assert np.array_equal(np.array(ve_p), ve_r, equal_nan=True), k
        if ion_r:
            n_ion += 1
            n_nan += int(np.isnan(vnew_r).any())
            assert np.array_equal(np.array(vnew_p), vnew_r, equal_nan=True), k
    assert n_ion > 40 and n_nan > 3                                   # both branches and the NaN product exercised
    # a whole apply() on a one-cell case (unordered_map with one key: the cell order is not an issue), E field on: products go through
    # Species::addParticle with its half-step rewind
    neu = util.random_particles(60, x0, 0.5 * xm, seed=31, vth=600.0, mpw=(5e11, 5e11)); neu[:, 6] = 5e11
    ele = _electrons_ev(30, x0, 0.5 * xm, 32, 8.0, 150.0)
    sn.setParticles(neu); se.setParticles(ele)
    sv_max = 8e-20 * 8e6
    m.setWsvMax(sv_max)
    ref2.seed(777); draws = [ref2.rnd() for _ in range(4000)]; ref2.seed(777)
    m.apply(dt)
    g = util.build_grid(orc, ni, nj, nk, x0, xm)

    def adder(charge, mass):
        def add(pos, vel, mpw):
            out = g.add_particles(ef, charge, mass, dt, np.array([pos + vel + [mpw]]))
            return [float(x) for x in out[0]] if len(out) else None
        return add
    ln, le = [list(map(float, r)) for r in neu], [list(map(float, r)) for r in ele]
    cand, coll, n_ion, ions, new_e, step_max = M.apply_cell(iter(draws), ln, le, dt, sv_max, adder(util.QE, 16 * util.AMU), adder(-util.QE, util.ME))
    assert cand > 30 and coll > 5 and n_ion > 0
    assert np.array_equal(sn.getParticles(), neu)                     # neutrals are never touched (:631)
    assert np.array_equal(se.getParticles(), np.array(le + new_e), equal_nan=True)
    assert np.array_equal(si.getParticles().reshape(-1, 7), np.array(ions).reshape(-1, 7))
    assert m.getWsvMax() == step_max                                  # :638-640
    for o in (m, sn, si, se, w):
        o.close()


# ----------------------------------------------------------------------------------------------- (2) device, pair by pair
@pytest.mark.gpu
def test_v2_cross_sections_on_the_device(picgpu):
    import mcc_restatement as R
    x0, xm, rects = util.discharge_geometry(7, 7, 9)
    w = util.build_world(picgpu.World, 7, 7, 9, x0, xm, rects)
    sn, si, se = _species(picgpu, w)
    tE, tS = util.momentum_transfer_table()
    m = picgpu.MC_MEX_Ionization(sn, si, se, w, tE, tS)
    m.setVariant(1)
    M = _model(R, 1.0)
    q = np.concatenate([np.logspace(-4, 7, 300), tE, [5.0, 9.793, 12.0, 13.6, 13.618, 13.62]])
    sc, sion = m.sigma(q)
    assert np.array_equal(sc, [M.sigma_coll(float(e)) for e in q])
    assert np.allclose(sion, [M.sigma_ion(float(e)) for e in q], rtol=1e-13, atol=1e-40)      # no threshold guard: negative below 9.793 eV
    assert sion[np.searchsorted(q[:300], 5.0)] < 0
    m.setVariant(0)
    assert m.sigma(np.array([12.0]))[1][0] == 0.0                    # the v3 guard is back
    for o in (m, sn, si, se, w):
        o.close()


@pytest.mark.gpu
def test_v2_one_cell_pair_by_pair(picgpu, orc):
    """One cell, E field on; the device's Philox stream replayed on the host drives MccModelV2 (pinned above against the compiled
    reference): same candidates, collisions, ionisations, products in the same order (NaN products included)."""
    import mcc_restatement as R
    from test_gpu_dsmc import _philox_stream
    pg = picgpu
    ni = nj = nk = 3
    x0, xm = np.array([0.0, 0.0, 0.0]), np.array([2e-3, 2e-3, 2e-3])
    seed, dt, sv_max = 4321, 1.8e-10, 8e-20 * 8e6
    ef = util.smooth_ef((ni, nj, nk), x0, xm, seed=5, amp=3e6)
    neu = util.random_particles(60, x0, 0.5 * xm, seed=31, vth=600.0, mpw=(5e11, 5e11)); neu[:, 6] = 5e11
    ele = _electrons_ev(30, x0, 0.5 * xm, 32, 8.0, 150.0)
    w = util.build_world(pg.World, ni, nj, nk, x0, xm, dt=dt)
    w.upload(pg.F_EF, ef)
    sn, si, se = _species(pg, w)
    sn.setParticles(neu); se.setParticles(ele)
    pg.seed(seed)
    tE, tS = util.momentum_transfer_table()
    m = pg.MC_MEX_Ionization(sn, si, se, w, tE, tS)
    m.setVariant(1); m.setWsvMax(sv_max)
    st = m.apply(dt)
    got_n, got_e, got_i = sn.getParticles(), se.getParticles(), si.getParticles()
    M = _model(R, 1e-9)
    g = util.build_grid(orc, ni, nj, nk, x0, xm)

    def adder(charge, mass):
        def add(pos, vel, mpw):
            out = g.add_particles(ef, charge, mass, dt, np.array([pos + vel + [mpw]]))
            return [float(x) for x in out[0]] if len(out) else None
        return add
    ln, le = [list(map(float, r)) for r in neu], [list(map(float, r)) for r in ele]
    stream = _philox_stream(orc, seed, 4 + 16 * 0, 0, 1)
    cand, coll, n_ion, ions, new_e, step_max = M.apply_cell(stream, ln, le, dt, sv_max, adder(util.QE, 16 * util.AMU), adder(-util.QE, util.ME))
    assert cand > 30 and coll > 5 and n_ion > 0
    assert (st.candidates, st.collisions, st.ionizations) == (cand, coll, n_ion)
    assert st.nan_products == sum(1 for q in new_e if np.isnan(q[3:6]).any())
    assert np.array_equal(got_n, neu)
    want_e = np.array(le + new_e)
    assert got_e.shape == want_e.shape and len(got_i) == len(ions)
    assert np.allclose(got_e, want_e, rtol=1e-11, atol=1e-6, equal_nan=True)
    assert np.array_equal(got_e[:, [0, 1, 2, 6]], want_e[:, [0, 1, 2, 6]])
    assert np.allclose(got_i, np.array(ions), rtol=1e-13, atol=0)     # ions: the neutral's velocity rewound by addParticle (no libm involved)
    assert abs(st.w_sigma_v_max - step_max) <= 1e-12 * step_max
    for o in (m, sn, si, se, w):
        o.close()


# ----------------------------------------------------------------------------------------------- (3) ensembles
def _agree(a, b, name, rel_floor=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
    diff = abs(a.mean() - b.mean())
    assert diff <= CI_SIGMA * se + rel_floor * abs(b.mean()), f"{name}: gpu {a.mean():.6g} vs ref {b.mean():.6g}, diff {diff:.3g} > {CI_SIGMA} * {se:.3g}"


def _energy_ev(parts):
    return 0.5 * util.ME * (parts[:, 3:6] ** 2).sum(1) / util.QE


def _run(mod, is_ref, seed, table, dt, sv_max, steps=2):
    ni, nj, nk = 7, 7, 9
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    rng = np.random.default_rng(1000)
    neu = util.random_particles(6000, x0, xm, seed=2000, vth=600.0, mpw=(5e11, 5e11), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    ele = util.random_particles(3000, x0, xm, seed=3000, vth=2.5e6, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    ele[:, 3:6] *= rng.uniform(0.5, 2.0, (len(ele), 1))
    w = util.build_world(mod.World, ni, nj, nk, x0, xm, rects, dt=dt)
    ef = util.smooth_ef((ni, nj, nk), x0, xm, seed=9, amp=2e5)
    if is_ref:
        w.setEF(ef)
    else:
        w.upload(mod.F_EF, ef)
    sn, si, se = _species(mod, w, (5e11, 50.0, 100.0))               # ions_to_create = 2 (:623)
    sn.setParticles(neu); se.setParticles(ele)
    mod.seed(seed)
    if is_ref:
        m = mod.MC_MEX_Ionization(sn, si, se, w, table)
    else:
        E, s = util.momentum_transfer_table()
        m = mod.MC_MEX_Ionization(sn, si, se, w, E, s); m.setVariant(1)
    m.setWsvMax(sv_max)
    for _ in range(steps):                                            # the second call runs with the ceiling sampled by the first (:638-640)
        m.apply(dt)
    pe, pi_, pn = se.getParticles(), si.getParticles(), sn.getParticles()
    ok = ~np.isnan(pe[:, 3:6]).any(1)
    # the device keeps its stores cell-sorted, so the created electrons are not simply the tail: they sit exactly on a neutral's position
    at_neutral = {tuple(r) for r in neu[:, 0:3]}
    new = np.array([tuple(r) in at_neutral for r in pe[:, 0:3]])
    out = dict(n_ion=len(pi_), n_ele=len(pe) - len(ele), n_new=int(new.sum()), n_nan=int((~ok).sum()), e_mean=_energy_ev(pe[ok]).mean(),
               e_new=_energy_ev(pe[new & ok]).mean() if (new & ok).any() else 0.0,
               w_ion=pi_[:, 6].sum() if len(pi_) else 0.0, uz_ion=pi_[:, 5].mean() if len(pi_) else 0.0, ion_z=pi_[:, 2].mean() if len(pi_) else 0.0,
               neutrals_untouched=bool(np.array_equal(util.sort_rows(pn), util.sort_rows(neu))))
    for o in (m, sn, si, se, w):
        o.close()
    return out


@pytest.mark.gpu
@pytest.mark.reference
def test_v2_ensemble(picgpu, ref2):
    dt = 1e-10
    sv_max = 8e-20 * 8e6
    with tempfile.TemporaryDirectory() as d:
        table = util.write_table(os.path.join(d, "Oxygen_momentum_transfer.txt"))
        G = [_run(picgpu, False, s, table, dt, sv_max) for s in range(N_SEEDS)]
        Rr = [_run(ref2, True, 100 + s, table, dt, sv_max) for s in range(N_SEEDS)]
    assert np.mean([r["n_ion"] for r in Rr]) > 40
    for key in ("n_ion", "n_ele", "n_nan", "e_mean", "e_new", "w_ion", "uz_ion", "ion_z"):
        _agree([g[key] for g in G], [r[key] for r in Rr], key)
    for g in G + Rr:
        assert g["neutrals_untouched"]                                # neutrals are never depleted in ch4/v2 (:631)
        assert g["n_ion"] == 2 * g["n_ele"] and g["n_new"] == g["n_ele"]   # two ions of mpw0 50 per ionisation, one electron of 100
