"""Slab-decomposed Poisson over peer memory (needs >= 2 GPUs on one node; skipped elsewhere, run with gpurun --gpus 2).

Every rank solves the same problem twice: replicated (the single-GPU red-black SOR on its own full grid) and slab-decomposed
(planes of i split over the ranks, halos / residual sum / all-gather through CUDA IPC peer memory inside the kernels).
The bar is bit-exactness: same iteration count, same residual, same phi on every rank."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "engineering-degree-in-plasma-simulations_b200"
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir, shape):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pg = importlib.import_module(PKG + ".picgpu")
    pg.init(rank)
    ni, nj, nk = shape
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    w = util.build_world(pg.World, ni, nj, nk, x0, xm, rects, dt=1e-10)
    rng = np.random.default_rng(11)                                    # identical on every rank
    rho = rng.normal(0.0, 2e-4, (ni, nj, nk))
    phi0 = rng.normal(0.0, 0.5, (ni, nj, nk))

    def all_gather_bytes(b):
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [bytes(o.cpu().numpy().tobytes()) for o in out]

    res = {}
    for mode in ("replicated", "slab"):
        for n0, tag in ((0.0, "lin"), (1e12, "boltz")):
            w.upload(pg.F_RHO, rho); w.upload(pg.F_PHI, phi0)
            sol = pg.PotentialSolver(w, 120, 1e-3)
            sol.setReferenceValues(0.0, n0, 1.5)
            if mode == "slab":
                sol.enableSlabs(rank, world, all_gather_bytes)
            conv = sol.solveGS()
            sol.computeEF()
            res[(mode, tag)] = (conv, sol.iterations, sol.L2, w.download(pg.F_PHI).copy(), w.download(pg.F_EF).copy())
            if mode == "slab":                                          # a second solve on the same solver: warm start, sequence numbers continue
                w.upload(pg.F_RHO, 0.5 * rho)
                sol.solveGS()
                res[(mode, tag + "2")] = w.download(pg.F_PHI).copy()
            else:
                w.upload(pg.F_RHO, 0.5 * rho)
                sol.solveGS()
                res[(mode, tag + "2")] = w.download(pg.F_PHI).copy()
            dist.barrier()
            sol.close()
    checks = []
    for tag in ("lin", "boltz"):
        a, b = res[("replicated", tag)], res[("slab", tag)]
        checks += [a[0] == b[0], a[1] == b[1], abs(a[2] - b[2]) <= 1e-12 * abs(a[2]),      # the residual is summed in a different order
                   np.array_equal(a[3], b[3]) and np.isfinite(a[3]).all(), np.array_equal(a[4], b[4]), np.array_equal(res[("replicated", tag + "2")], res[("slab", tag + "2")])]
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([int(all(checks)), res[("slab", "lin")][1]] + [int(c) for c in checks]))
    np.save(os.path.join(out_dir, f"phi_{rank}.npy"), res[("slab", "lin")][3])
    w.close()
    dist.destroy_process_group()


def _density_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pg = importlib.import_module(PKG + ".picgpu"); mg = importlib.import_module(PKG + ".multigpu")
    pg.init(rank)
    ni, nj, nk = 4 * world, 9, 13                                      # planes divisible by the ranks: equal reduce-scatter chunks
    x0, xm, rects = util.discharge_geometry(ni, nj, nk)
    w = util.build_world(pg.World, ni, nj, nk, x0, xm, rects, dt=1e-10)
    parts = util.random_particles(60001, x0, xm, seed=19, mpw=(1.0, 5e3))          # identical on every rank
    lo = sum(mg.split_count(len(parts), r, world) for r in range(rank))
    mine = parts[lo:lo + mg.split_count(len(parts), rank, world)]
    ion_all = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0); ele_all = pg.Species("e-", util.ME, -util.QE, w, 100.0)
    ion = pg.Species("O+", 16 * util.AMU, util.QE, w, 100.0); ele = pg.Species("e-", util.ME, -util.QE, w, 100.0)
    S = 30
    for sp, p in ((ion_all, parts), (ele_all, parts[::-1]), (ion, mine), (ele, mine[::-1])):
        sp.setParticles(p); sp.sort(); sp.setDensityScale(S)
    ion_all.computeNumberDensity(); ele_all.computeNumberDensity()
    w.computeChargeDensity([ion_all, ele_all])
    den_ref, rho_ref = ion_all.den.copy(), w.rho.copy()
    nv = ni * nj * nk
    rng_ = (rank * (nv // world), (rank + 1) * (nv // world))
    for sp in (ion, ele):
        sp.depositPartial()
        v = mg.fixed_view(torch, sp, pg.SF_DEN_FIXED)
        with torch.cuda.stream(torch.cuda.ExternalStream(pg.stream_ptr(), device=rank)):
            dist.reduce_scatter_tensor(list(v.chunk(world))[rank], v)             # in place: the r-th chunk of the own input
        sp.finalizeDensity(rng_)
    w.upload(pg.F_RHO, np.zeros((ni, nj, nk)))
    w.computeChargeDensity([ion, ele], rng_)
    ok = np.array_equal(ion.den.reshape(-1)[rng_[0]:rng_[1]], den_ref.reshape(-1)[rng_[0]:rng_[1]])
    ok &= np.array_equal(w.rho.reshape(-1)[rng_[0]:rng_[1]], rho_ref.reshape(-1)[rng_[0]:rng_[1]])
    ok &= not w.rho.reshape(-1)[:rng_[0]].any() and not w.rho.reshape(-1)[rng_[1]:].any()      # nothing written outside the owned planes
    np.save(os.path.join(out_dir, f"dok_{rank}.npy"), np.array([int(ok)]))
    dist.barrier(); dist.destroy_process_group()


def test_reduce_scatter_onto_slabs_matches_single_gpu_density(tmp_path):
    """Particles split by index over the ranks, int64 accumulators reduce-scattered in place onto the Poisson slabs, density and
    charge density finalised on the owned planes only: bit-identical there to the single-GPU deposit of all particles."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs on the node")
    mp.spawn(_density_worker, args=(world, 29800 + (os.getpid() % 2000), str(tmp_path)), nprocs=world, join=True)
    assert all(np.load(tmp_path / f"dok_{r}.npy")[0] == 1 for r in range(world))


@pytest.mark.parametrize("shape", [(12, 9, 13), (33, 16, 20)])
def test_slab_solve_is_bit_identical_to_the_replicated_solve(tmp_path, shape):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs on the node")
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path), shape), nprocs=world, join=True)
    oks = [np.load(tmp_path / f"ok_{r}.npy") for r in range(world)]
    assert all(o[0] == 1 for o in oks), [[int(v) for v in o] for o in oks]
    assert oks[0][1] > 25                                               # several residual checks were exchanged
    phis = [np.load(tmp_path / f"phi_{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert np.array_equal(phis[0], phis[r])                         # every rank ends with the same full phi
