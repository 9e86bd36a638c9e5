"""NR-PCG on the device (ch4/v3/src/PotentialSolver.cpp:178-347) against the compiled reference's solve().

The reference's matrix is not symmetric (NEUMANN rows), its CG usually diverges and solveGSlinear finishes the Newton step;
both sides therefore agree on converged potentials, not on iterates: 1e-6 relative, the north-star bound for phi.
"""
import numpy as np
import pytest

import util

pytestmark = [pytest.mark.gpu, pytest.mark.reference]


def _cubic_geometry(n, V):
    L = 0.008
    x0, xm = np.array([-0.004, -0.004, 0.0]), np.array([0.004, 0.004, L])
    rects = [((0.0, 0.0, x0[2]), V, (L, L, 0.1 * L)), ((0.0, 0.0, xm[2]), -V, (L, L, 0.1 * L))]
    return x0, xm, rects


@pytest.mark.parametrize("case", ["cubic-linear", "noncubic-boltzmann"])
def test_nrpcg_matches_reference_solve(picgpu, ref, case):
    if case == "cubic-linear":
        ni, nj, nk = 9, 9, 9
        x0, xm, rects = _cubic_geometry(9, -4000.0)
        n0, Te0, amp = 0.0, 1e20, 1e-6
    else:                                                   # dx != dz: the reference's PCG matrix has the x/z swap
        ni, nj, nk = 11, 9, 13
        x0, xm, rects = util.discharge_geometry(ni, nj, nk, phi=-2.0)
        n0, Te0, amp = 1e12, 1.5, 1e-9
    rho = amp * np.random.default_rng(3).normal(size=(ni, nj, nk))
    wr = util.build_world(ref.World, ni, nj, nk, x0, xm, rects); wr.set(1, rho)
    sr = ref.PotentialSolver(wr, 5000, 1e-4, ref.PotentialSolver.PCG); sr.setReferenceValues(0.0, n0, Te0)
    assert sr.solve()
    want = wr.get(0)
    wg = util.build_world(picgpu.World, ni, nj, nk, x0, xm, rects); wg.upload(picgpu.F_RHO, rho)
    sg = picgpu.PotentialSolver(wg, 5000, 1e-4); sg.setReferenceValues(0.0, n0, Te0)
    assert sg.solveNRPCG(xz_swap=True)
    assert util.norm_err(wg.phi, want) < 1e-6
    if case != "cubic-linear":
        # without the swap the device solves the equation solveGS relaxes: it must then agree with the device GS, and differ from the reference's PCG
        wg2 = util.build_world(picgpu.World, ni, nj, nk, x0, xm, rects); wg2.upload(picgpu.F_RHO, rho)
        s2 = picgpu.PotentialSolver(wg2, 5000, 1e-4); s2.setReferenceValues(0.0, n0, Te0)
        assert s2.solveNRPCG(xz_swap=False)
        wg3 = util.build_world(picgpu.World, ni, nj, nk, x0, xm, rects); wg3.upload(picgpu.F_RHO, rho)
        s3 = picgpu.PotentialSolver(wg3, 40000, 1e-7); s3.setReferenceValues(0.0, n0, Te0)
        assert s3.solveGS()
        assert util.norm_err(wg2.phi, wg3.phi) < 1e-6
        assert util.norm_err(wg2.phi, want) > 1e-3
        for o in (s2, wg2, s3, wg3):
            o.close()
    for o in (sr, wr, sg, wg):
        o.close()
