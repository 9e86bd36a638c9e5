"""Parity of the CUDA path (through the C ABI, libpicgpu.so) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): deposition bit-exact (int64 fixed point); potential, E field and pushed
particle state <= 1e-6 relative in fp64 (the kernels are written FMA-free, so most comparisons are in
fact bit-for-bit and assert that).  Where oracle/_ref is present the same outputs are also compared with
the compiled reference itself.
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

TOL = 1e-6   # fp64 relative tolerance stated by north_star


def _setup(picgpu, orc, ni=11, nj=9, nk=13, spheres=(), dt=1e-12):
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    w = util.build_world(picgpu.World, ni, nj, nk, x0, xm, rects, spheres, dt=dt)
    g = util.build_grid(orc, ni, nj, nk, x0, xm, rects, spheres)
    return w, g, x0, xm


def test_world_geometry_matches_oracle(picgpu, orc):
    w, g, x0, xm = _setup(picgpu, orc, 41, 41, 61)
    assert np.array_equal(w.node_vol, g.node_volumes())
    oid, phi = g.compute_object_id()
    assert np.array_equal(w.object_id, oid.astype(float))
    assert np.array_equal(w.phi, phi)
    w.close()


def test_upload_download_roundtrip(picgpu, orc):
    w, g, x0, xm = _setup(picgpu, orc)
    parts = util.random_particles(10007, x0, xm, seed=1)
    sp = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    sp.setParticles(parts)
    assert sp.getNumParticles() == len(parts)
    assert np.array_equal(sp.getParticles(), parts)
    ef = util.smooth_ef((w.ni, w.nj, w.nk), x0, xm)
    w.upload(picgpu.F_EF, ef)
    assert np.array_equal(w.ef, ef)
    sp.setParticles(np.empty((0, 7)))
    assert sp.getNumParticles() == 0
    sp.close(); w.close()


@pytest.mark.parametrize("n", [1, 31, 5000, 200000])
def test_push_electrons_matches_oracle(picgpu, orc, n):
    sph = [((0.0, 0.0, 0.0025), -100.0, 0.0007)]
    w, g, x0, xm = _setup(picgpu, orc, spheres=sph)
    ef = util.smooth_ef((w.ni, w.nj, w.nk), x0, xm, seed=3, amp=3e6)
    w.upload(picgpu.F_EF, ef)
    parts = util.random_particles(n, x0, xm, seed=4 + n, vth=2e6)
    sp = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    sp.setParticles(parts)
    dt = 2e-10
    sp.advanceElectrons(dt)
    want, alive = g.push_electrons(ef, -util.QE, util.ME, dt, parts)
    got = sp.getParticles()
    assert sp.getNumParticles() == alive.sum()
    a, b = util.sort_rows(got), util.sort_rows(want[alive])
    assert util.rel_err(a, b) <= TOL
    assert np.array_equal(a, b)          # FMA-free kernels: bit for bit
    sp.close(); w.close()


@pytest.mark.parametrize("n,steps", [(8000, 1), (24000, 3), (50000, 2), (90000, 2), (400000, 1)])
def test_push_of_a_sorted_store_matches_oracle(picgpu, orc, n, steps):
    """Electron push of a cell-sorted store against the oracle, bit for bit, over several pushes on an increasingly stale
    partition (stragglers, holes, deaths) and over a range of cell populations."""
    sph = [((0.0, 0.0, 0.0025), -100.0, 0.0007)]
    w, g, x0, xm = _setup(picgpu, orc, spheres=sph)
    ef = util.smooth_ef((w.ni, w.nj, w.nk), x0, xm, seed=5, amp=3e6)
    w.upload(picgpu.F_EF, ef)
    parts = util.random_particles(n, x0, xm, seed=40 + n, vth=2e6)
    sp = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    sp.setParticles(parts); sp.sort()
    cur = sp.getParticles()
    for it in range(steps):
        sp.advanceElectrons(5e-11)
        want, alive = g.push_electrons(ef, -util.QE, util.ME, 5e-11, cur)
        got = sp.getParticles()
        assert sp.getNumParticles() == alive.sum() and alive.sum() < len(cur)
        assert np.array_equal(util.sort_rows(got), util.sort_rows(want[alive]))
        cur = got
    sp.close(); w.close()


def test_push_electrons_everything_dies_and_nothing_dies(picgpu, orc):
    w, g, x0, xm = _setup(picgpu, orc)
    ef = np.zeros((w.ni, w.nj, w.nk, 3))
    w.upload(picgpu.F_EF, ef)
    parts = util.random_particles(3000, x0, xm, seed=11, vth=1.0, lo_frac=(0, 0, 0.3), hi_frac=(1, 1, 0.7))
    sp = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    sp.setParticles(parts)
    sp.advanceElectrons(1e-12)
    assert sp.getNumParticles() == 3000
    fast = parts.copy(); fast[:, 5] = 1e9
    sp.setParticles(fast)
    sp.advanceElectrons(1e-6)
    assert sp.getNumParticles() == 0
    sp.advanceElectrons(1e-6)            # empty store
    assert sp.getNumParticles() == 0
    sp.close(); w.close()


def test_push_reflect_matches_oracle(picgpu, orc):
    x0, xm = np.array([-0.1, -0.1, 0.0]), np.array([0.1, 0.1, 0.2])
    w = util.build_world(picgpu.World, 21, 21, 21, x0, xm, dt=2e-10)
    g = util.build_grid(orc, 21, 21, 21, x0, xm)
    ef = util.smooth_ef((21, 21, 21), x0, xm, seed=5, amp=50.0)
    w.upload(picgpu.F_EF, ef)
    parts = util.random_particles(50000, x0, xm, seed=6, vth=3e7)
    sp = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    sp.setParticles(parts)
    sp.advanceReflect(2e-10)
    want = g.push_reflect(ef, -util.QE, util.ME, 2e-10, parts)
    assert np.any(want[:, 3:6] * parts[:, 3:6] < 0)     # some reflections happened
    assert np.array_equal(sp.getParticles(), want)
    sp.close(); w.close()


def test_add_particles_filter_and_rewind(picgpu, orc):
    w, g, x0, xm = _setup(picgpu, orc)
    ef = util.smooth_ef((w.ni, w.nj, w.nk), x0, xm, seed=5, amp=1e6)
    w.upload(picgpu.F_EF, ef)
    parts = util.random_particles(4000, x0 - 0.1 * (xm - x0), xm + 0.1 * (xm - x0), seed=6)
    parts[7, 3] = np.nan; parts[9, 1] = np.nan
    sp = picgpu.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
    acc = sp.addParticles(parts)
    want = g.add_particles(ef, util.QE, 16 * util.AMU, 1e-12, parts)
    assert acc == len(want) and 0 < acc < len(parts)
    assert np.array_equal(util.sort_rows(sp.getParticles()), util.sort_rows(want))
    acc2 = sp.addParticles(parts[:100])                  # appends after existing particles
    assert sp.getNumParticles() == acc + acc2
    sp.close(); w.close()


@pytest.mark.parametrize("n,sort", [(1, False), (777, False), (60000, False), (60000, True), (400000, True)])
def test_deposit_bit_exact(picgpu, orc, n, sort):
    w, g, x0, xm = _setup(picgpu, orc)
    parts = util.random_particles(n, x0, xm, seed=7 + n, mpw=(1.0, 5e11))
    sp = picgpu.Species("O", 16 * util.AMU, 0.0, w, 5e11)
    sp.setParticles(parts)
    if sort:
        sp.sort()
    sp.computeNumberDensity()
    S = sp.densityScale()
    fixed = g.deposit_fixed(parts, S)
    assert np.array_equal(sp.den_fixed, fixed)                                   # bit-exact
    vol = g.node_volumes()
    assert np.array_equal(sp.den, g.finalize_density(fixed, S, vol))
    assert util.norm_err(sp.den, g.deposit_fp64(parts, vol)) < 1e-12             # vs the reference's fp64 sum, normwise
    assert fixed.max() < 2 ** 62 and fixed.max() >= 2 ** 45                      # scale calibration left headroom and resolution
    # deterministic: a second deposit gives the same bits
    sp.computeNumberDensity()
    assert np.array_equal(sp.den_fixed, fixed)
    sp.close(); w.close()


@pytest.mark.parametrize("n,lo,hi", [(50000, (0.3, 0.0, 0.45), (0.6, 1.0, 0.55)), (300000, (0.0, 0.5, 0.12), (1.0, 0.62, 0.88)), (40000, (0.0, 0.0, 0.12), (1.0, 1.0, 0.88)), (20000, (0.0, 0.0, 0.12), (1.0, 1.0, 0.88))])
def test_deposit_cell_partition_ragged(picgpu, orc, n, lo, hi):
    """Cell-partition deposit on ragged cell populations: empty passes, cells holding many chunks, count-only pass, and the
    same bits as the thread-run kernel (unsorted store) and the oracle."""
    w, g, x0, xm = _setup(picgpu, orc)
    parts = util.random_particles(n, x0, xm, seed=91 + n, mpw=(1.0, 1e6), lo_frac=lo, hi_frac=hi)
    a = picgpu.Species("O", 16 * util.AMU, 0.0, w, 1.0)
    b = picgpu.Species("O", 16 * util.AMU, 0.0, w, 1.0)
    a.setParticles(parts); a.sort(); a.setDensityScale(24)
    b.setParticles(parts); b.setDensityScale(24)            # unsorted: generic kernel
    a.computeNumberDensity(); b.computeNumberDensity()
    fixed = g.deposit_fixed(parts, 24)
    assert np.array_equal(a.den_fixed, fixed) and np.array_equal(b.den_fixed, fixed)
    cnt = g.count_per_cell(parts)
    a.computeMacroParticlesCount()                           # by-product of the deposit pass
    assert np.array_equal(a.macro_part_count, cnt)
    a.advanceNonElectron(a, a, 0.0)                          # a (null) push invalidates the cached count: count-only pass (all particles lie in the gap)
    a.computeMacroParticlesCount()
    assert np.array_equal(a.macro_part_count, cnt)
    a.close(); b.close(); w.close()


def test_deposit_pinned_scale_and_empty(picgpu, orc):
    w, g, x0, xm = _setup(picgpu, orc)
    sp = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    sp.setDensityScale(20)
    sp.computeNumberDensity()
    assert not sp.den_fixed.any() and not sp.den.any()
    parts = util.random_particles(5000, x0, xm, seed=17)
    sp.setParticles(parts)
    sp.computeNumberDensity()
    assert sp.densityScale() == 20
    assert np.array_equal(sp.den_fixed, g.deposit_fixed(parts, 20))
    sp.close(); w.close()


def test_fused_push_deposit_equals_separate(picgpu, orc):
    w, g, x0, xm = _setup(picgpu, orc)
    ef = util.smooth_ef((w.ni, w.nj, w.nk), x0, xm, seed=3, amp=3e6)
    w.upload(picgpu.F_EF, ef)
    parts = util.random_particles(80000, x0, xm, seed=21, vth=2e6)
    a = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    b = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    for sp in (a, b):
        sp.setParticles(parts); sp.sort(); sp.setDensityScale(30)
    a.advanceElectrons(2e-10); a.computeNumberDensity(); a.computeMacroParticlesCount()
    b.advanceElectronsDeposit(2e-10, count_cells=True)
    assert a.getNumParticles() == b.getNumParticles()
    assert np.array_equal(a.den_fixed, b.den_fixed)
    assert np.array_equal(a.den, b.den)
    assert np.array_equal(a.macro_part_count, b.macro_part_count)
    assert np.array_equal(util.sort_rows(a.getParticles()), util.sort_rows(b.getParticles()))
    want, alive = g.push_electrons(ef, -util.QE, util.ME, 2e-10, parts)
    assert np.array_equal(b.den_fixed, g.deposit_fixed(want[alive], 30))
    a.close(); b.close(); w.close()


def test_fused_heavy_push_deposit_equals_separate(picgpu, orc):
    sph = [((0.0, 0.0, 0.0025), -100.0, 0.0007)]
    w, g, x0, xm = _setup(picgpu, orc, spheres=sph, dt=1e-9)
    ef = util.smooth_ef((w.ni, w.nj, w.nk), x0, xm, seed=3, amp=2e5)
    w.upload(picgpu.F_EF, ef)
    parts = util.random_particles(90000, x0, xm, seed=23, vth=4e4, mpw=(100.0, 100.0), lo_frac=(0, 0, 0.1), hi_frac=(1, 1, 0.9))
    neu = picgpu.Species("O", 16 * util.AMU, 0.0, w, 5e11)
    a = picgpu.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
    b = picgpu.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
    for sp in (a, b):
        sp.setParticles(parts); sp.sort(); sp.setDensityScale(30)
    a.advanceNonElectron(neu, neu, 2e-8); a.computeNumberDensity(); a.computeMacroParticlesCount()
    b.advanceNonElectronDeposit(neu, neu, 2e-8, count_cells=True)
    assert 0 < a.getNumParticles() < len(parts) and a.getNumParticles() == b.getNumParticles()
    assert np.array_equal(a.den_fixed, b.den_fixed)
    assert np.array_equal(a.macro_part_count, b.macro_part_count)
    assert np.array_equal(util.sort_rows(a.getParticles()), util.sort_rows(b.getParticles()))
    assert np.array_equal(a.den_fixed, g.deposit_fixed(a.getParticles(), 30))
    # partial (multi-GPU) path: raw accumulator, finalised separately
    c = picgpu.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
    c.setParticles(parts); c.sort(); c.setDensityScale(30)
    c.advanceDepositPartial(2e-8, neu, neu, heavy=True)
    assert np.array_equal(c.den_fixed, a.den_fixed)
    c.finalizeDensity()
    assert np.array_equal(c.den, a.den)
    for o in (a, b, c, neu, w):
        o.close()


def test_cell_partition_path_with_stragglers_holes_and_appended_tail(picgpu, orc):
    """The warp-per-cell fast path must not depend on how stale the partition is: after the sort particles drift out of
    their cells, die (holes are filled from the tail) and new ones are appended beyond the partition."""
    w, g, x0, xm = _setup(picgpu, orc)
    ef = util.smooth_ef((w.ni, w.nj, w.nk), x0, xm, seed=3, amp=3e6)
    w.upload(picgpu.F_EF, ef)
    parts = util.random_particles(70000, x0, xm, seed=31, vth=1.5e6)
    sp = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    sp.setParticles(parts); sp.sort(); sp.setDensityScale(28)
    cur = sp.getParticles()
    for it in range(4):                                   # several pushes on one (increasingly stale) partition
        if it % 2 == 0:                                   # push, then deposit + count over the cell partition (lane group per cell)
            sp.advanceElectrons(6e-11); sp.computeNumberDensity(); sp.computeMacroParticlesCount()
        else:                                             # fused push + deposit (thread runs)
            sp.advanceElectronsDeposit(6e-11, count_cells=True)
        want, alive = g.push_electrons(ef, -util.QE, util.ME, 6e-11, cur)
        cur = want[alive]
        assert sp.getNumParticles() == len(cur)
        assert np.array_equal(sp.den_fixed, g.deposit_fixed(cur, 28))
        assert np.array_equal(sp.macro_part_count, g.count_per_cell(cur))
        got = sp.getParticles()
        assert np.array_equal(util.sort_rows(got), util.sort_rows(cur))
        cur = got                                         # follow the device order (compaction permutes)
        if it == 1:                                       # append beyond the partition
            extra = util.random_particles(9000, x0, xm, seed=32, vth=1.5e6, lo_frac=(0, 0, 0.2), hi_frac=(1, 1, 0.8))
            sp.addParticles(extra)
            cur = sp.getParticles()
    assert len(cur) < 70000 + 9000
    sp.computeNumberDensity()                             # deposit-only through the same path
    assert np.array_equal(sp.den_fixed, g.deposit_fixed(cur, 28))
    sp.close(); w.close()


def test_count_per_cell_and_sort(picgpu, orc):
    w, g, x0, xm = _setup(picgpu, orc)
    parts = util.random_particles(123457, x0, xm, seed=8)
    sp = picgpu.Species("e-", util.ME, -util.QE, w, 100.0)
    sp.setParticles(parts)
    sp.computeMacroParticlesCount()
    cnt = g.count_per_cell(parts)
    assert np.array_equal(sp.macro_part_count, cnt)
    sp.sort()
    got = sp.getParticles()
    assert np.array_equal(util.sort_rows(got), util.sort_rows(parts))      # a permutation of the input
    lc = (got[:, :3] - x0) * (1.0 / ((xm - x0) / (np.array([w.ni, w.nj, w.nk]) - 1)))
    ijk = lc.astype(np.int64)
    cell = (ijk[:, 0] * (w.nj - 1) + ijk[:, 1]) * (w.nk - 1) + ijk[:, 2]
    assert np.all(np.diff(cell) >= 0)                                       # sortedness
    sp.computeMacroParticlesCount()
    assert np.array_equal(sp.macro_part_count, cnt)                         # idempotent under permutation
    # stable: within a cell the original relative order is kept
    first_cell = cell[0]
    same = got[cell == first_cell]
    orig_lc = ((parts[:, :3] - x0) * (1.0 / ((xm - x0) / (np.array([w.ni, w.nj, w.nk]) - 1)))).astype(np.int64)
    orig_cell = (orig_lc[:, 0] * (w.nj - 1) + orig_lc[:, 1]) * (w.nk - 1) + orig_lc[:, 2]
    assert np.array_equal(same, parts[orig_cell == first_cell])
    sp.close(); w.close()


def test_charge_density(picgpu, orc):
    w, g, x0, xm = _setup(picgpu, orc)
    sps, dens, qs = [], [], []
    for i, (m, q) in enumerate([(16 * util.AMU, 0.0), (16 * util.AMU, util.QE), (util.ME, -util.QE)]):
        parts = util.random_particles(3000, x0, xm, seed=20 + i)
        sp = picgpu.Species("s%d" % i, m, q, w, 1.0)
        sp.setParticles(parts); sp.computeNumberDensity()
        sps.append(sp); dens.append(sp.den); qs.append(q)
    w.computeChargeDensity(sps)
    assert np.array_equal(w.rho, g.charge_density(dens, qs))
    for sp in sps:
        sp.close()
    w.close()


@pytest.mark.parametrize("n0,Te0", [(0.0, 1e20), (1.5, 1e10), (1e12, 5000.0)])
def test_poisson_fixed_point_and_ef(picgpu, orc, n0, Te0):
    w, g, x0, xm = _setup(picgpu, orc, 13, 11, 17)
    rng = np.random.default_rng(9)
    rho = rng.normal(0, 1e-7, (13, 11, 17))
    w.upload(picgpu.F_RHO, rho)
    oid, phi_start = g.compute_object_id()
    sol = picgpu.PotentialSolver(w, 20000, 1e-4)
    sol.setReferenceValues(0.0, n0, Te0)
    assert sol.solveGS()
    phi_gpu = w.phi
    # (1) same iteration (red-black) on the CPU: same iterates up to exp() ulps
    phi_rb, conv, its, l2 = g.solve_rb(oid, rho, phi_start, 20000, 1e-4, 0.0, n0, Te0)
    assert conv and its == sol.iterations
    assert util.norm_err(phi_gpu, phi_rb) <= 1e-12
    # (2) the reference's lexicographic Gauss-Seidel, converged: shared fixed point within north_star's 1e-6
    phi_gs, conv, _, _ = g.solve_gs(oid, rho, phi_start, 20000, 1e-4, 0.0, n0, Te0)
    assert conv
    assert util.norm_err(phi_gpu, phi_gs) <= TOL
    assert abs(sol.residual() - g.residual(oid, rho, phi_gpu, 0.0, n0, Te0)) <= 1e-9 * max(1.0, sol.L2)
    sol.computeEF()
    assert np.array_equal(w.ef, g.compute_ef(phi_gpu))     # E from identical phi: bit for bit
    sol.close(); w.close()


def test_poisson_ch2_dirichlet_box(picgpu, orc):
    x0, xm = np.array([-0.1, -0.1, 0.0]), np.array([0.1, 0.1, 0.2])
    w = util.build_world(picgpu.World, 21, 21, 21, x0, xm)
    g = util.build_grid(orc, 21, 21, 21, x0, xm)
    rng = np.random.default_rng(12)
    rho = rng.normal(0, 1e-9, (21, 21, 21))
    w.upload(picgpu.F_RHO, rho)
    sol = picgpu.PotentialSolver(w, 10000, 1e-4)
    sol.setBoundaryMode(1)
    assert sol.solveGS()
    oid = np.zeros((21, 21, 21), dtype=np.int32)
    phi_gs, conv, _, _ = g.solve_gs(oid, rho, np.zeros((21, 21, 21)), 10000, 1e-4, 0.0, 0.0, 1e20, bc_mode=1)
    assert conv
    phi = w.phi
    assert not phi[0].any() and not phi[:, 0].any() and not phi[:, :, -1].any()     # faces stay 0 (Dirichlet box)
    assert util.norm_err(phi, phi_gs) <= 1e-4   # both stopped at L2 < 1e-4 from different sweep orders
    sol.close(); w.close()


def test_moments_and_diagnostics(picgpu, orc):
    w, g, x0, xm = _setup(picgpu, orc)
    parts = util.random_particles(20000, x0, xm, seed=30)
    sp = picgpu.Species("O+", 16 * util.AMU, util.QE, w, 100.0)
    sp.setParticles(parts)
    sp.sampleMoments(); sp.sampleMoments()
    sums = g.sample_moments(parts); sums = g.sample_moments(parts, sums)
    for fid, s in zip((picgpu.SF_N_SUM, picgpu.SF_NV_SUM, picgpu.SF_NUU_SUM, picgpu.SF_NVV_SUM, picgpu.SF_NWW_SUM), sums):
        assert util.norm_err(sp.download(fid), s) < 1e-12
    mc, mom, ke = sp.diagnostics()
    m = parts[:, 6]
    assert abs(mc - m.sum()) / m.sum() < 1e-12
    assert np.allclose(mom, 16 * util.AMU * (m[:, None] * parts[:, 3:6]).sum(0), rtol=1e-10, atol=1e-30)
    assert abs(ke - 0.5 * 16 * util.AMU * (m * (parts[:, 3:6] ** 2).sum(1)).sum()) / ke < 1e-12
    sp.clearSamples()
    assert not sp.download(picgpu.SF_N_SUM).any()
    sp.close(); w.close()


def test_error_reporting(picgpu):
    with pytest.raises(picgpu.PicgError):
        picgpu.World(2, 2, 2, (0, 0, 0), (1, 1, 1))
    w = picgpu.World(5, 5, 5, (0, 0, 0), (1, 1, 1))
    with pytest.raises(picgpu.PicgError):
        picgpu.Species("x", -1.0, 0.0, w, 1.0)
    w.close()


@pytest.mark.reference
def test_against_compiled_reference(picgpu, ref):
    """End-to-end on the device vs the unmodified reference: push + deposit + charge density + E field."""
    ni, nj, nk = 11, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    wr = util.build_world(ref.World, ni, nj, nk, x0, xm, rects)
    wg = util.build_world(picgpu.World, ni, nj, nk, x0, xm, rects)
    rng = np.random.default_rng(40)
    phi = wr.get(0) + np.where(wr.get(4) > 0, 0.0, rng.normal(0, 50.0, wr.shape))
    wr.set(0, phi); wg.upload(picgpu.F_PHI, phi)
    sr = ref.PotentialSolver(wr, 10, 1e-4, ref.PotentialSolver.GS); sr.computeEF()
    sg = picgpu.PotentialSolver(wg, 10, 1e-4); sg.computeEF()
    assert np.array_equal(wg.ef, wr.get(3))
    parts = util.random_particles(30000, x0, xm, seed=41, vth=2e6)
    er = ref.Species("e-", util.ME, -util.QE, wr, 100.0); er.setParticles(parts)
    eg = picgpu.Species("e-", util.ME, -util.QE, wg, 100.0); eg.setParticles(parts)
    er.advanceElectrons(5e-11); eg.advanceElectrons(5e-11)
    assert np.array_equal(util.sort_rows(eg.getParticles()), util.sort_rows(er.getParticles()))
    er.computeNumberDensity(); eg.computeNumberDensity()
    assert util.norm_err(eg.den, er.get(0)) < 1e-12
    er.computeMacroParticlesCount(); eg.computeMacroParticlesCount()
    assert np.array_equal(eg.macro_part_count, er.get(4))
    wr.computeChargeDensity([er]); wg.computeChargeDensity([eg])
    assert util.norm_err(wg.rho, wr.get(1)) < 1e-12
    for o in (er, sr, wr, eg, sg, wg):
        o.close()


def test_neutral_push_is_a_pure_drift_and_matches_the_reference(picgpu, ref):
    """charge == 0: the device skips the E gather and the velocity write-back (k_run<.., DRIFT>); the reference adds
    E*(dt*0/m) = 0 (Species.cpp:187-188).  Same particles bit for bit, in a non-zero field, with deaths at the box faces."""
    ni, nj, nk = 11, 9, 13
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    wr = util.build_world(ref.World, ni, nj, nk, x0, xm, rects)
    wg = util.build_world(picgpu.World, ni, nj, nk, x0, xm, rects)
    ef = util.smooth_ef((ni, nj, nk), x0, xm, seed=5, amp=3e6)
    wr.set(3, ef); wg.upload(picgpu.F_EF, ef)
    # inside the gap, too slow to reach an electrode in one step (no stochastic re-emission), fast enough to leave through x/y
    parts = util.random_particles(50000, x0, xm, seed=43, vth=3e3, mpw=(5e11, 5e11), lo_frac=(0, 0, 0.35), hi_frac=(1, 1, 0.65))
    nr = ref.Species("O", 16 * util.AMU, 0.0, wr, 5e11); nr.setParticles(parts)
    ng = picgpu.Species("O", 16 * util.AMU, 0.0, wg, 5e11); ng.setParticles(parts)
    picgpu.timers_reset(); picgpu.timers_enable(True)
    for _ in range(2):
        nr.advanceNonElectron(nr, nr, 4e-8); ng.advanceNonElectron(ng, ng, 4e-8)
    picgpu.timers_enable(False)
    assert "push_neutral" in picgpu.timers_read() and "push_heavy" not in picgpu.timers_read()      # the drift kernel is what ran
    got, want = util.sort_rows(ng.getParticles()), util.sort_rows(nr.getParticles())
    assert 0 < len(want) < len(parts)
    assert np.array_equal(got, want)
    for o in (nr, wr, ng, wg):
        o.close()


@pytest.mark.parametrize("shape,n0,Te0,bc", [((13, 11, 17), 0.0, 1e20, 0), ((13, 11, 17), 1e12, 5000.0, 0), ((37, 21, 133), 0.0, 1e20, 0), ((9, 40, 70), 0.0, 1e20, 1), ((5, 5, 5), 0.0, 1e20, 0)])
def test_tiled_one_pass_sweep_is_bit_identical_to_the_row_sweeps(picgpu, shape, n0, Te0, bc):
    """PotentialSolver::solveGS as ONE plane-marching shared-memory pass per iteration (k_sor_tiled, both colours, phi double-buffered)
    against the two row sweeps per iteration (k_sor_row): the same bits after any number of iterations (odd and even batches, tiles
    and plane chunks that do not divide the mesh, electrodes, Boltzmann electrons, the ch2 Dirichlet box), the same convergence decision."""
    pg = picgpu
    ni, nj, nk = shape
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    rng = np.random.default_rng(21)
    rho = rng.normal(0, 1e-7, shape)
    phis = {}
    for mode in (0, 1):
        w = util.build_world(pg.World, ni, nj, nk, x0, xm, rects if bc == 0 else ())
        w.upload(pg.F_RHO, rho)
        sol = pg.PotentialSolver(w, 400, 1e-3)
        sol.setReferenceValues(0.0, n0, Te0); sol.setBoundaryMode(bc); sol.setSweep(mode)
        out = []
        for n in (1, 2, 7, 25, 26):                           # plain launches (n < 4) and graph replays, odd and even batch lengths
            sol.iterate(n); out.append(w.phi)
        conv = sol.solveGS(); out.append(w.phi)
        sol.computeEF(); out.append(w.ef)
        phis[mode] = (out, conv, sol.iterations, sol.L2)
        sol.close(); w.close()
    for a, b in zip(phis[0][0], phis[1][0]):
        assert np.array_equal(a, b)
    assert phis[0][1:] == phis[1][1:]
