"""DSMC_MEX (ch4/v3/src/Interactions.cpp:143-285) on the device against the compiled reference.

The pair selection is stochastic (Philox streams here, the global mt19937 there), so the comparison is the ensemble test
of tests/test_gpu_stochastic.py: N_SEEDS seeds on each side on identical inputs, two-sample z test at CI_SIGMA combined
standard errors.  What does not involve the RNG is compared exactly: the cross-section, the number of candidate pairs
(a function of the per-cell counts only), conservation of momentum and energy by every run, the constructor's error.
"""
import numpy as np
import pytest

import util
from test_gpu_stochastic import CI_SIGMA, N_SEEDS, _agree

pytestmark = [pytest.mark.gpu, pytest.mark.reference]

NI, NJ, NK = 7, 7, 9
MPW0 = 1e13
DT = 2e-9
SV_MAX = 2e-15            # a realistic ceiling of sigma*v_rel for O at ~1 km/s: about one candidate in four collides


def _case(n, seed, vth_xyz):
    x0, xm, _ = util.discharge_geometry(NI, NJ, NK)
    p = util.random_particles(n, x0, xm, seed=seed, vth=1.0, mpw=(MPW0, MPW0))
    p[:, 3:6] *= np.asarray(vth_xyz)
    p[:, 6] = MPW0
    return x0, xm, p


def _moments(p, mass):
    return dict(sxx=(p[:, 3] ** 2).sum(), syy=(p[:, 4] ** 2).sum(), szz=(p[:, 5] ** 2).sum(), mom=mass * p[:, 3:6].sum(0),
                ke=0.5 * mass * (p[:, 3:6] ** 2).sum())


def _by_position(p):
    """Collisions change velocities only; the device may reorder the store (cell sort), so rows are matched by position."""
    return p[np.lexsort((p[:, 2], p[:, 1], p[:, 0]))]


def _run_one(mod, seed, n_apply):
    x0, xm, p0 = _case(8000, 11, (1500.0, 300.0, 300.0))          # hot along x: collisions relax the anisotropy
    w = util.build_world(mod.World, NI, NJ, NK, x0, xm, dt=DT)
    sp = mod.Species("O", 16 * util.AMU, 0.0, w, MPW0)
    sp.setParticles(p0)
    mod.seed(seed)
    m = mod.DSMC_MEX(sp, w)
    m.setSigmaVMax(SV_MAX)
    cand = []
    for _ in range(n_apply):
        st = m.apply(DT)
        if st is not None:
            cand.append((st.candidates, st.collisions))
    p1 = _by_position(sp.getParticles()); p0 = _by_position(p0)
    out = _moments(p1, 16 * util.AMU)
    out["changed"] = int((p1[:, 3:6] != p0[:, 3:6]).any(1).sum())
    out["unchanged_rest"] = bool(np.array_equal(p1[:, [0, 1, 2, 6]], p0[:, [0, 1, 2, 6]]))
    out["sv_max"] = m.getSigmaVMax() if hasattr(m, "getSigmaVMax") else m.stats.sigma_v_max
    out["cand"] = cand
    out["p0"] = p0
    for o in (m, sp, w):
        o.close()
    return out


def test_sigma_matches_reference(picgpu, ref):
    x0, xm, p0 = _case(10, 1, (1, 1, 1))
    wg = util.build_world(picgpu.World, NI, NJ, NK, x0, xm); wr = util.build_world(ref.World, NI, NJ, NK, x0, xm)
    sg = picgpu.Species("O", 16 * util.AMU, 0.0, wg, MPW0); sr = ref.Species("O", 16 * util.AMU, 0.0, wr, MPW0)
    mg, mr = picgpu.DSMC_MEX(sg, wg), ref.DSMC_MEX(sr, wr)
    v = np.exp(np.random.default_rng(5).uniform(np.log(1e-2), np.log(1e7), 4000))
    assert util.rel_err(mg.sigma(v), mr.sigma(v)) < 1e-14           # pow() on the device: within 2 ulp of libm's
    for o in (mg, mr, sg, sr, wg, wr):
        o.close()


def test_constructor_rejects_unequal_weights(picgpu, ref):
    x0, xm, _ = _case(10, 1, (1, 1, 1))
    wg = util.build_world(picgpu.World, NI, NJ, NK, x0, xm); wr = util.build_world(ref.World, NI, NJ, NK, x0, xm)
    a = picgpu.Species("O", 16 * util.AMU, 0.0, wg, 1e13); b = picgpu.Species("O2", 32 * util.AMU, 0.0, wg, 2e13)
    with pytest.raises(picgpu.PicgError) as e:
        picgpu.DSMC_MEX(a, b, wg)
    assert "same macroparticle weight" in str(e.value)              # the reference's message (Interactions.cpp:161)
    ra = ref.Species("O", 16 * util.AMU, 0.0, wr, 1e13); rb = ref.Species("O2", 32 * util.AMU, 0.0, wr, 2e13)
    with pytest.raises(ValueError):
        ref.DSMC_MEX(ra, rb, wr)
    for o in (a, b, ra, rb, wg, wr):
        o.close()


def test_one_species_candidates_and_conservation(picgpu):
    """No RNG involved: the candidate count of the first call is sum_c int(0.5*np*np*mpw0*sv*dt/dV + 0.5) (:196-197);
    every collision conserves momentum and energy; positions and weights are untouched."""
    g = _run_one(picgpu, 3, 1)
    x0, xm, _ = util.discharge_geometry(NI, NJ, NK)
    dx = (xm - x0) / (np.array([NI, NJ, NK]) - 1)
    ijk = np.minimum(((g["p0"][:, 0:3] - x0) * (1.0 / dx)).astype(int), [NI - 2, NJ - 2, NK - 2])
    counts = np.bincount((ijk[:, 0] * (NJ - 1) + ijk[:, 1]) * (NK - 1) + ijk[:, 2], minlength=(NI - 1) * (NJ - 1) * (NK - 1)).astype(float)
    dv = dx[0] * dx[1] * dx[2]
    want = ((0.5 * counts * counts * MPW0 * SV_MAX * DT / dv + 0.5).astype(int) * (counts >= 2)).sum()
    assert g["cand"][0][0] == want
    assert 0 < g["cand"][0][1] < want
    m0 = _moments(g["p0"], 16 * util.AMU)
    assert abs(g["ke"] - m0["ke"]) <= 1e-12 * m0["ke"]
    assert np.abs(g["mom"] - m0["mom"]).max() <= 1e-12 * 16 * util.AMU * np.abs(g["p0"][:, 3:6]).sum()
    assert g["unchanged_rest"]
    assert g["changed"] <= 2 * g["cand"][0][1]


def test_one_species_ensemble(picgpu, ref):
    """Three applies per seed: the second and third run with the sigma_v_rel_max the previous call sampled (:219-222)."""
    G = [_run_one(picgpu, s, 3) for s in range(N_SEEDS)]
    R = [_run_one(ref, 100 + s, 3) for s in range(N_SEEDS)]
    m0 = _moments(G[0]["p0"], 16 * util.AMU)
    assert np.mean([r["sxx"] for r in R]) < 0.97 * m0["sxx"]         # the case relaxes visibly
    for key in ("sxx", "syy", "szz", "changed", "sv_max"):
        _agree([g[key] for g in G], [r[key] for r in R], key)
    for run in G + R:
        assert abs(run["ke"] - m0["ke"]) <= 1e-11 * m0["ke"]
        assert run["unchanged_rest"]


def _run_two(mod, seed, n_apply):
    x0, xm, pa = _case(6000, 21, (1200.0, 1200.0, 1200.0))
    _, _, pb = _case(5000, 22, (250.0, 250.0, 250.0))
    pb[:, 3] += 400.0                                                # species 2 drifts: momentum flows to species 1
    w = util.build_world(mod.World, NI, NJ, NK, x0, xm, dt=DT)
    a = mod.Species("O", 16 * util.AMU, 0.0, w, MPW0); b = mod.Species("O2", 32 * util.AMU, 0.0, w, MPW0)
    a.setParticles(pa); b.setParticles(pb)
    mod.seed(seed)
    m = mod.DSMC_MEX(a, b, w)
    m.setSigmaVMax(SV_MAX)
    for _ in range(n_apply):
        m.apply(DT)
    qa, qb = _by_position(a.getParticles()), _by_position(b.getParticles())
    pa, pb = _by_position(pa), _by_position(pb)
    ma, mb = _moments(qa, 16 * util.AMU), _moments(qb, 32 * util.AMU)
    out = dict(ke_a=ma["ke"], ke_b=mb["ke"], px_a=ma["mom"][0], px_b=mb["mom"][0], mom=ma["mom"] + mb["mom"], ke=ma["ke"] + mb["ke"],
               changed_a=int((qa[:, 3:6] != pa[:, 3:6]).any(1).sum()), changed_b=int((qb[:, 3:6] != pb[:, 3:6]).any(1).sum()),
               ke0=_moments(pa, 16 * util.AMU)["ke"] + _moments(pb, 32 * util.AMU)["ke"],
               mom0=_moments(pa, 16 * util.AMU)["mom"] + _moments(pb, 32 * util.AMU)["mom"],
               pabs=16 * util.AMU * np.abs(pa[:, 3:6]).sum() + 32 * util.AMU * np.abs(pb[:, 3:6]).sum())
    for o in (m, a, b, w):
        o.close()
    return out


def test_two_species_ensemble(picgpu, ref):
    G = [_run_two(picgpu, s, 2) for s in range(N_SEEDS)]
    R = [_run_two(ref, 100 + s, 2) for s in range(N_SEEDS)]
    assert np.mean([r["changed_b"] for r in R]) > 200
    for key in ("ke_a", "ke_b", "px_a", "px_b", "changed_a", "changed_b"):
        _agree([g[key] for g in G], [r[key] for r in R], key)
    for run in G:
        assert abs(run["ke"] - run["ke0"]) <= 1e-12 * run["ke0"]
        assert np.abs(run["mom"] - run["mom0"]).max() <= 1e-12 * run["pabs"]


def test_collisions_on_a_stale_partition_equal_a_fresh_sort(picgpu):
    """The per-cell lists come from the mover machinery when the partition is stale: same seed, same lists (as sets) -> the
    candidate count is identical with and without a forced re-sort, and the collision count agrees statistically."""
    pg = picgpu
    x0, xm, p0 = _case(20000, 31, (900.0, 900.0, 900.0))
    res = []
    for frac in (0.5, 0.0):                                         # 0.5: patch the stale partition; 0.0: always re-sort
        pg.set_mover_fraction(frac); pg.seed(9)
        w = util.build_world(pg.World, NI, NJ, NK, x0, xm, dt=DT)
        sp = pg.Species("O", 16 * util.AMU, 0.0, w, MPW0)
        sp.setParticles(p0); sp.sort()
        sp.advanceNonElectron(sp, sp, 2e-7)                         # ~10 % of the particles change cell, some leave the box
        m = pg.DSMC_MEX(sp, w); m.setSigmaVMax(SV_MAX)
        st = m.apply(DT)
        res.append((st.candidates, st.collisions, sp.getNumParticles()))
        for o in (m, sp, w):
            o.close()
    pg.set_mover_fraction(0.1)
    assert res[0][0] == res[1][0] and res[0][2] == res[1][2]
    assert abs(res[0][1] - res[1][1]) < 6 * np.sqrt(res[1][1])


def _philox_stream(orc, seed, stream_id, index, step):
    """The device's PhiloxStream (csrc/philox.cuh): key = (lo32(seed) ^ stream, hi32(seed)), counter = (index lo, index hi, step, block);
    a block yields two doubles of 53 bits, the one from words 3:2 first."""
    k = [(seed & 0xffffffff) ^ stream_id, seed >> 32]
    block = 0
    while True:
        out = orc.philox4x32([index & 0xffffffff, index >> 32, step, block], k)
        block += 1
        yield float(((int(out[3]) << 32) | int(out[2])) >> 11) * 2.0 ** -53
        yield float(((int(out[1]) << 32) | int(out[0])) >> 11) * 2.0 ** -53


def test_one_cell_pair_by_pair_against_the_restatement(picgpu, orc):
    """Deterministic check below the ensemble level: one cell, the device's random stream replayed on the host (Philox restated in the
    oracle), the reference's candidate loop (Interactions.cpp:185-223) run with the oracle's evaluateSigma / collide.  Same pairs, same
    collisions; velocities agree to the last digits (device sin / cos / sqrt are within an ulp or two of libm's)."""
    pg = picgpu
    x0, xm = np.array([0.0, 0.0, 0.0]), np.array([2e-3, 2e-3, 2e-3])            # 3 nodes per axis (the minimum): 8 cells of 1 mm, all particles in cell 0
    n, seed, dt, sv_max, mass = 40, 4242, 3e-9, 2e-15, 16 * util.AMU
    p = util.random_particles(n, x0, 0.5 * xm, seed=3, vth=900.0, mpw=(MPW0, MPW0)); p[:, 6] = MPW0
    w = util.build_world(pg.World, 3, 3, 3, x0, xm, dt=dt)
    sp = pg.Species("O", mass, 0.0, w, MPW0)
    sp.setParticles(p)
    pg.seed(seed)
    m = pg.DSMC_MEX(sp, w); m.setSigmaVMax(sv_max)
    st = m.apply(dt)
    got = sp.getParticles()
    assert np.array_equal(got[:, [0, 1, 2, 6]], p[:, [0, 1, 2, 6]])             # one cell: the stable sort keeps the order
    # host replay
    r = _philox_stream(orc, seed, 6 + 16 * 0, 0, 1)                             # RNG_DSMC = 6, species 0, rank 0; cell 0, first call
    v = p[:, 3:6].copy()
    dv = 1e-3 * 1e-3 * 1e-3
    n_groups = int(0.5 * n * n * MPW0 * sv_max * dt / dv + 0.5)
    n_coll, sv_seen = 0, 0.0
    for _ in range(n_groups):
        a = int(next(r) * n); b = int(next(r) * n)
        while a == b:
            b = int(next(r) * n)
        d = v[a] - v[b]
        v_rel = float(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        sv = orc.dsmc_sigma(mass, mass, v_rel) * v_rel
        sv_seen = max(sv_seen, sv)
        if sv / sv_max > next(r):
            n_coll += 1
            v[a], v[b] = orc.dsmc_collide(mass, mass, next(r), next(r), v[a], v[b])
    assert n_groups > 30 and 0 < n_coll < n_groups
    assert (st.candidates, st.collisions) == (n_groups, n_coll)
    assert np.allclose(got[:, 3:6], v, rtol=1e-12, atol=1e-9)
    assert abs(st.sigma_v_max - sv_seen) <= 1e-13 * sv_seen                     # the ceiling of the next call (:219-222)
    for o in (m, sp, w):
        o.close()


def test_two_species_one_cell_pair_by_pair(picgpu, orc):
    """applyTwoSpecies (Interactions.cpp:225-265) replayed draw by draw in one cell, like the one-species case above."""
    pg = picgpu
    x0, xm = np.array([0.0, 0.0, 0.0]), np.array([2e-3, 2e-3, 2e-3])
    na, nb, seed, dt, sv_max = 30, 24, 99, 3e-9, 2e-15
    ma, mb = 16 * util.AMU, 32 * util.AMU
    pa = util.random_particles(na, x0, 0.5 * xm, seed=4, vth=1100.0, mpw=(MPW0, MPW0)); pa[:, 6] = MPW0
    pb = util.random_particles(nb, x0, 0.5 * xm, seed=5, vth=300.0, mpw=(MPW0, MPW0)); pb[:, 6] = MPW0
    w = util.build_world(pg.World, 3, 3, 3, x0, xm, dt=dt)
    a = pg.Species("O", ma, 0.0, w, MPW0); b = pg.Species("O2", mb, 0.0, w, MPW0)
    a.setParticles(pa); b.setParticles(pb)
    pg.seed(seed)
    m = pg.DSMC_MEX(a, b, w); m.setSigmaVMax(sv_max)
    st = m.apply(dt)
    ga, gb = a.getParticles(), b.getParticles()
    r = _philox_stream(orc, seed, 6 + 16 * 0, 0, 1)                             # the stream of species1 (index 0 in this world)
    va, vb = pa[:, 3:6].copy(), pb[:, 3:6].copy()
    n_groups = int(0.5 * na * nb * MPW0 * sv_max * dt / 1e-9 + 0.5)
    n_coll = 0
    for _ in range(n_groups):
        i = int(next(r) * na); j = int(next(r) * nb)
        d = va[i] - vb[j]
        v_rel = float(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        if orc.dsmc_sigma(ma, mb, v_rel) * v_rel / sv_max > next(r):
            n_coll += 1
            va[i], vb[j] = orc.dsmc_collide(ma, mb, next(r), next(r), va[i], vb[j])
    assert n_groups > 15 and 0 < n_coll < n_groups
    assert (st.candidates, st.collisions) == (n_groups, n_coll)
    assert np.allclose(ga[:, 3:6], va, rtol=1e-12, atol=1e-9) and np.allclose(gb[:, 3:6], vb, rtol=1e-12, atol=1e-9)
    for o in (m, a, b, w):
        o.close()
