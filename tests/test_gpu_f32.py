"""The fp32 secondary path (csrc/f32.cu): cell-relative single-precision particle store.

north_star: "pushed particle state must match to within 1e-6 relative for fp64 and 1e-4 relative for fp32"; "deposition must be bit-exact
on a fixed particle set".  Checked here:
  * push against the fp64 oracle and against the reference itself built with `using type_calc = float` (oracle/_ref/libref_v3_f32.so):
    velocities and displacements within 1e-4 relative, the same particles absorbed (but for the few that end within rounding of a face);
  * deposit: the int64 fixed-point grid bit for bit against a numpy float32 restatement of Field::scatter (Field.h:157-199) on the
    stored cell-relative coordinates; the density within 1e-6 (normwise) of the fp64 oracle; per-cell counts exact;
  * store round trip, cell sort (multiset kept, cells ascending, same deposit bits), several steps against the fp64 device path.
"""
import numpy as np
import pytest

import util

pytestmark = [pytest.mark.gpu]

NI, NJ, NK = 13, 11, 17
TOL32 = 1e-4            # the north star's fp32 tolerance


def _case(n=60000, seed=3, vth=1.2e6):
    x0, xm, rects = util.discharge_geometry(NI, NJ, NK)
    p = util.random_particles(n, x0, xm, seed=seed, vth=vth, mpw=(50.0, 200.0), lo_frac=(0, 0, 0.12), hi_frac=(1, 1, 0.88))
    ef = util.smooth_ef((NI, NJ, NK), x0, xm, seed=5, amp=4e5)
    return x0, xm, rects, p, ef


def _stored_fractions(p, x0, xm):
    """Cell indices and float32 fractions exactly as the store forms them (f32.cu: to_cell_frac)."""
    n = np.array([NI, NJ, NK])
    inv_dx = 1.0 / ((xm - x0) / (n - 1))
    l = (p[:, 0:3] - x0) * inv_dx
    c = np.clip(l.astype(np.int64), 0, n - 2)
    f = (l - c).astype(np.float32)
    f = np.where(f >= np.float32(1.0), np.float32(0.99999994), f)
    return c, f


def test_store_round_trip_and_sort(picgpu):
    pg = picgpu
    x0, xm, rects, p, ef = _case()
    w = util.build_world(pg.World, NI, NJ, NK, x0, xm, rects)
    s = pg.Species32("e-", util.ME, -util.QE, w, 100.0)
    s.setParticles(p)
    got = s.getParticles()
    dx = (xm - x0) / (np.array([NI, NJ, NK]) - 1)
    assert got.shape == p.shape
    assert np.all(np.abs(got[:, 0:3] - p[:, 0:3]) <= dx * 2.0 ** -23)                 # half a float32 spacing of a fraction in [0, 1)
    assert np.array_equal(got[:, 3:7], p[:, 3:7].astype(np.float32).astype(np.float64))
    s.sort()
    srt = s.getParticles()
    assert np.array_equal(util.sort_rows(srt), util.sort_rows(got))
    c, _ = _stored_fractions(srt, x0, xm)
    key = (c[:, 0] * (NJ - 1) + c[:, 1]) * (NK - 1) + c[:, 2]
    assert np.all(np.diff(key) >= 0)
    for o in (s, w):
        o.close()


def _compare_push(got, want, alive_got_n, p, dt, x0, xm):
    """got: pushed survivors of the fp32 path (any order); want / alive: the fp64 result for every input particle (input order)."""
    pushed, alive = want
    # match survivors through their weights (unique in these cases) - the device compacts, the order changes
    order_w = {float(np.float32(m)): i for i, m in enumerate(p[:, 6])}
    idx = np.array([order_w[m] for m in got[:, 6]])
    assert len(set(idx)) == len(idx)
    L = xm - x0
    margin = np.min(np.minimum(pushed[:, 0:3] - x0, xm - pushed[:, 0:3]) / L, axis=1)
    disagree = np.setxor1d(idx, np.nonzero(alive)[0])
    assert len(disagree) <= 5 and np.all(np.abs(margin[disagree]) < 1e-5), (len(disagree), margin[disagree])    # only particles that end on a face
    both = np.isin(idx, np.nonzero(alive)[0])
    g, r, src = got[both], pushed[idx[both]], p[idx[both]]
    dv = np.linalg.norm(g[:, 3:6] - r[:, 3:6], axis=1) / np.linalg.norm(r[:, 3:6], axis=1)
    assert dv.max() < TOL32, dv.max()
    disp = r[:, 0:3] - src[:, 0:3]
    dd = np.linalg.norm((g[:, 0:3] - src[:, 0:3]) - disp, axis=1) / np.linalg.norm(disp, axis=1)
    assert np.median(dd) < 1e-6 and dd.max() < TOL32, (np.median(dd), dd.max())
    return len(g)


def test_push_matches_fp64_oracle_and_fp32_reference(picgpu, orc):
    pg = picgpu
    x0, xm, rects, p, ef = _case()
    p[:, 6] = 50.0 + np.arange(len(p)) * 0.25                                        # unique weights, exact in float32: the particle's identity
    dt = 2e-11
    w = util.build_world(pg.World, NI, NJ, NK, x0, xm, rects, dt=dt)
    w.upload(pg.F_EF, ef)
    s = pg.Species32("e-", util.ME, -util.QE, w, 100.0)
    s.setParticles(p)
    start = s.getParticles()                                                         # the positions the store actually holds (rounded fractions)
    s.advanceElectrons(dt)
    got = s.getParticles()
    g = util.build_grid(orc, NI, NJ, NK, x0, xm, rects)
    want = g.push_electrons(ef, -util.QE, util.ME, dt, start)
    n_cmp = _compare_push(got, want, len(got), start, dt, x0, xm)
    assert n_cmp > 0.8 * len(p) and (~want[1]).sum() > 100                           # most survive, some are absorbed (box and electrodes)
    # the reference itself in single precision (all.h:11 -> float)
    from oracle import ref_v3_f32 as r32
    if not r32.available():
        pytest.skip("oracle/_ref/libref_v3_f32.so not built")
    r32.lib(); r32.config(subcycling=False, multithreading=False, merging=False, sputtering=False)
    wr = util.build_world(r32.World, NI, NJ, NK, x0, xm, rects, dt=dt)
    wr.set(3, ef)
    sr = r32.Species("e-", util.ME, -util.QE, wr, 100.0)
    sr.setParticles(start)
    sr.advanceElectrons(dt)
    ref = sr.getParticles()
    # the reference keeps ABSOLUTE float positions: compare velocities at 1e-4 and positions at the float spacing of the coordinate
    wmap = {float(m): i for i, m in enumerate(ref[:, 6])}
    common = [(k, wmap[m]) for k, m in enumerate(got[:, 6]) if m in wmap]
    assert len(common) > 0.98 * min(len(got), len(ref))
    a = got[[k for k, _ in common]]; b = ref[[j for _, j in common]]
    dv = np.linalg.norm(a[:, 3:6] - b[:, 3:6], axis=1) / np.linalg.norm(b[:, 3:6], axis=1)
    assert dv.max() < TOL32, dv.max()
    assert np.all(np.abs(a[:, 0:3] - b[:, 0:3]) <= 4 * np.spacing(np.float32(np.abs(b[:, 0:3]).max())))
    for o in (s, w, sr, wr):
        o.close()


def test_deposit_bit_exact_against_float32_restatement(picgpu, orc):
    pg = picgpu
    x0, xm, rects, p, ef = _case(n=80000)
    w = util.build_world(pg.World, NI, NJ, NK, x0, xm, rects)
    s = pg.Species32("O+", 16 * util.AMU, util.QE, w, 100.0)
    s.setParticles(p)
    s.computeNumberDensity()
    S = s.densityScale()
    c, f = _stored_fractions(p, x0, xm)
    m = p[:, 6].astype(np.float32)
    one = np.float32(1.0)
    di, dj, dk = f[:, 0], f[:, 1], f[:, 2]
    odi, odj, odk = one - di, one - dj, one - dk
    vs = m * np.float32(2.0 ** S)
    want = np.zeros(NI * NJ * NK, dtype=np.int64)
    for a, wi in ((0, odi), (1, di)):
        for b, wj in ((0, odj), (1, dj)):
            for d, wk in ((0, odk), (1, dk)):
                contrib = np.rint(((vs * wi) * wj) * wk).astype(np.int64)             # ((val*wi)*wj)*wk in float32, then to the fixed-point grid
                node = ((c[:, 0] + a) * NJ + (c[:, 1] + b)) * NK + (c[:, 2] + d)
                np.add.at(want, node, contrib)
    assert np.array_equal(s.den_fixed.ravel(), want)
    g = util.build_grid(orc, NI, NJ, NK, x0, xm, rects)
    den64 = g.deposit_fp64(p, g.node_volumes())
    assert util.norm_err(s.den, den64) < 1e-6
    assert np.array_equal(s.macro_part_count, g.count_per_cell(p))
    s.sort(); s.computeNumberDensity()                                               # integer sums: any particle order gives the same bits
    assert np.array_equal(s.den_fixed.ravel(), want)
    # charge density of fp32 species = sum of charge * den
    e = pg.Species32("e-", util.ME, -util.QE, w, 100.0)
    e.setParticles(p[:30000]); e.computeNumberDensity()
    pg.charge_density32(w, [s, e])
    assert np.array_equal(w.rho, util.QE * s.den + (-util.QE) * e.den)
    for o in (s, e, w):
        o.close()


def test_several_steps_against_the_fp64_device_path(picgpu):
    pg = picgpu
    x0, xm, rects, p, ef = _case(n=100000, vth=6e5)
    dt = 1e-11
    w = util.build_world(pg.World, NI, NJ, NK, x0, xm, rects, dt=dt)
    w.upload(pg.F_EF, ef)
    a = pg.Species("e-", util.ME, -util.QE, w, 100.0); a.setParticles(p)
    b = pg.Species32("e-", util.ME, -util.QE, w, 100.0); b.fromSpecies(a)               # device-side conversion
    for _ in range(8):
        a.advanceElectrons(dt); b.advanceElectrons(dt)
    a.computeNumberDensity(); b.computeNumberDensity()
    na, nb = a.getNumParticles(), b.getNumParticles()
    assert abs(na - nb) <= 10 and na < 0.995 * len(p)                                   # some are absorbed on the way
    ka, kb = a.diagnostics()[2], b.diagnostics()[2]
    assert abs(ka - kb) / ka < TOL32
    assert util.norm_err(b.den, a.den) < 1e-3                                           # a handful of particles may sit in different cells / be absorbed
    for o in (a, b, w):
        o.close()
