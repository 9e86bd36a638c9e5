"""Species::merge on the device against the compiled reference (oracle/_ref) on identical particles.

The merged pairs (weights and velocities) are deterministic given the particle order: bit-exact multisets are required.  Only the two
member positions that receive the pair are random (mt19937 there, Philox here): they must be positions of original particles."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _clustered(n, x0, xm, seed, mpw):
    """Particles whose velocities fall into a few tight clusters per region, so that many share a velocity bin."""
    rng = np.random.default_rng(seed)
    a = util.random_particles(n, x0, xm, seed=seed, vth=1.0, mpw=mpw, lo_frac=(0, 0, 0.15), hi_frac=(1, 1, 0.85))
    centres = rng.normal(0.0, 4e4, (6, 3))
    k = rng.integers(0, 6, n)
    a[:, 3:6] = centres[k] + rng.normal(0.0, 40.0, (n, 3))
    return a


def _rows(a, cols):
    b = a[:, cols]
    return b[np.lexsort(b.T[::-1])]


@pytest.mark.parametrize("charge,ef_amp", [(0.0, 2e5), (1.0, 0.0)])
def test_merge_matches_reference(picgpu, ref, charge, ef_amp):
    ni, nj, nk = 7, 6, 9
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    w = util.build_world(picgpu.World, ni, nj, nk, x0, xm, rects, dt=1e-10)
    rw = util.build_world(ref.World, ni, nj, nk, x0, xm, rects, dt=1e-10)
    ef = util.smooth_ef((ni, nj, nk), x0, xm, seed=3, amp=ef_amp)
    w.upload(picgpu.F_EF, ef); rw.set(3, ef)
    parts = _clustered(30000, x0, xm, seed=5, mpw=(1.0, 50.0))
    sp = picgpu.Species("X", 16 * util.AMU, charge * util.QE, w, 10.0)
    sp.setParticles(parts); sp.sort()
    order = sp.getParticles()                                   # the device's cell order: the reference gets the same order
    rs = ref.Species("X", 16 * util.AMU, charge * util.QE, rw, 10.0)
    rs.setParticles(order)
    ref.seed(7)
    rs.merge()
    want = rs.getParticles()
    n0, n1, st = sp.merge()
    got = sp.getParticles()
    assert n0 == len(parts) and n1 == len(got) == len(want) and n1 < n0 // 2          # the clusters really merge
    assert st[0] > 1000 and st[1] == n0 - n1 and st[2] == 0 and st[3] == 0
    # weights and velocities of every particle (merged pairs and untouched ones): the same multiset, bit for bit
    assert np.array_equal(_rows(got, [3, 4, 5, 6]), _rows(want, [3, 4, 5, 6]))
    # every position is the position of an original particle
    orig = {tuple(r) for r in order[:, :3]}
    assert all(tuple(r) in orig for r in got[:, :3])
    # conservation: weight exactly (halves of sums), momentum and per-axis energy to rounding (E = 0 case: no half-step rewind)
    assert abs(got[:, 6].sum() - order[:, 6].sum()) <= 1e-9 * order[:, 6].sum()
    if ef_amp == 0.0:
        for c in (3, 4, 5):
            assert abs((got[:, 6] * got[:, c]).sum() - (order[:, 6] * order[:, c]).sum()) <= 1e-9 * np.abs(order[:, 6] * order[:, c]).sum()
            assert abs((got[:, 6] * got[:, c] ** 2).sum() - (order[:, 6] * order[:, c] ** 2).sum()) <= 1e-9 * (order[:, 6] * order[:, c] ** 2).sum()
    for o in (sp, w, rs, rw):
        o.close()


def test_merge_on_a_stale_partition_and_small_cells(picgpu):
    """After pushes the per-cell lists are patched with movers: the merge must still conserve weight / momentum / energy, never
    touch cells with fewer than 10 particles, and leave a store that deposits and sorts normally."""
    ni, nj, nk = 7, 6, 9
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    w = util.build_world(picgpu.World, ni, nj, nk, x0, xm, rects, dt=1e-10)
    parts = _clustered(20000, x0, xm, seed=9, mpw=(5.0, 5.0))
    sp = picgpu.Species("O", 16 * util.AMU, 0.0, w, 5.0)
    sp.setParticles(parts); sp.sort()
    sp.advanceNonElectron(sp, sp, 2e-9); sp.advanceNonElectron(sp, sp, 2e-9)      # drift: stragglers, a few deaths
    before = sp.getParticles()
    n0, n1, st = sp.merge()
    after = sp.getParticles()
    assert n0 == len(before) and n1 == len(after) < n0
    assert abs(after[:, 6].sum() - before[:, 6].sum()) <= 1e-9 * before[:, 6].sum()
    for c in (3, 4, 5):
        assert abs((after[:, 6] * after[:, c]).sum() - (before[:, 6] * before[:, c]).sum()) <= 1e-9 * np.abs(before[:, 6] * before[:, c]).sum()
        assert abs((after[:, 6] * after[:, c] ** 2).sum() - (before[:, 6] * before[:, c] ** 2).sum()) <= 1e-9 * (before[:, 6] * before[:, c] ** 2).sum()
    sp.computeNumberDensity(); sp.sort()
    assert sp.getNumParticles() == n1
    # a sparse species (fewer than 10 per cell everywhere) is left alone
    few = picgpu.Species("O", 16 * util.AMU, 0.0, w, 5.0)
    few.setParticles(_clustered(300, x0, xm, seed=10, mpw=(5.0, 5.0))); few.sort()
    m0, m1, st2 = few.merge()
    assert m0 == m1 == 300 and st2[0] == 0
    for o in (sp, few, w):
        o.close()
