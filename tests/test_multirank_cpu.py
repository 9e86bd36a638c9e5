"""The N>1 protocol on CPU: world_size-2 gloo processes run the rank-agreement rules of multigpu.py with the CPU oracle
standing in for the device kernels.  Checks: even split by index, a common fixed-point scale, and that the all-reduced
int64 accumulators are bit-identical to a single-rank deposit of all particles (the property the GPU path relies on)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "engineering-degree-in-plasma-simulations_b200"


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from oracle import pic_oracle as orc
    mg = importlib.import_module(PKG + ".multigpu")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x0, xm, rects = util.discharge_geometry(11, 9, 13)
    g = util.build_grid(orc, 11, 9, 13, x0, xm, rects)
    parts = util.random_particles(30001, x0, xm, seed=5, mpw=(1.0, 5e11))         # identical on every rank
    # even split by index
    lo = sum(mg.split_count(len(parts), r, world) for r in range(rank))
    mine = parts[lo:lo + mg.split_count(len(parts), rank, world)]
    # every rank calibrates on its own share; the common scale is the minimum minus ceil(log2 G)
    local_max = float(np.abs(g.deposit_fp64(mine, np.ones(g.shape))).max())
    S_local = 54 - int(np.floor(np.log2(local_max)))
    def reduce_min(v):
        t = torch.tensor([v], dtype=torch.int64); dist.all_reduce(t, op=dist.ReduceOp.MIN); return int(t.item())
    S = mg.common_scale(S_local, world, reduce_min)
    fixed = torch.from_numpy(g.deposit_fixed(mine, S).reshape(-1).copy())
    dist.all_reduce(fixed)                                                          # int64 sum
    np.save(os.path.join(out_dir, f"fixed_{rank}.npy"), fixed.numpy())
    np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([S, lo, len(mine)]))
    dist.destroy_process_group()


def test_two_rank_deposit_is_bit_identical_to_single_rank(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from oracle import pic_oracle as orc
    x0, xm, rects = util.discharge_geometry(11, 9, 13)
    g = util.build_grid(orc, 11, 9, 13, x0, xm, rects)
    parts = util.random_particles(30001, x0, xm, seed=5, mpw=(1.0, 5e11))
    f0, f1 = np.load(tmp_path / "fixed_0.npy"), np.load(tmp_path / "fixed_1.npy")
    m0, m1 = np.load(tmp_path / "meta_0.npy"), np.load(tmp_path / "meta_1.npy")
    assert m0[0] == m1[0]                                   # same scale on both ranks
    assert m0[2] + m1[2] == len(parts) and m0[1] == 0 and m1[1] == m0[2] and abs(m0[2] - m1[2]) <= 1
    assert np.array_equal(f0, f1)                           # all-reduce leaves identical grids everywhere
    S = int(m0[0])
    whole = g.deposit_fixed(parts, S).reshape(-1)
    assert np.array_equal(f0, whole)                        # and they equal the single-rank deposit, bit for bit
    assert whole.max() < 2 ** 62                            # headroom kept after summing G ranks


def test_rank_rules():
    mg = importlib.import_module(PKG + ".multigpu")
    assert [mg.split_count(10, r, 4) for r in range(4)] == [3, 3, 2, 2]
    assert sum(mg.split_count(10 ** 9 + 7, r, 8) for r in range(8)) == 10 ** 9 + 7
    assert mg.common_scale(20, 1, lambda v: v) == 20
    assert mg.common_scale(20, 8, lambda v: v - 1) == 16    # min over ranks (19) minus log2(8)
    assert mg.common_scale(20, 3, lambda v: v) == 18
    assert mg.common_ceiling(0.25, lambda v: max(v, 0.5)) == 0.5


def test_candidate_shares_add_up_to_the_single_rank_count():
    """The device rule of csrc/mcc.cu / dsmc.cu, restated in multigpu.candidate_share: with equal local populations the ranks'
    shares add up to the reference's rounded count on the whole cell, for any world size, cell and call number."""
    mg = importlib.import_module(PKG + ".multigpu")
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 4, 8):
        for frac_total in list(rng.uniform(0.0, 6.0, 40)) + [0.49, 0.5, 0.51, 1.5, 7.5]:
            frac_local = frac_total / (world * world)            # bilinear in the local counts: (np_n/G)(np_e/G)
            for cell, call in ((0, 1), (5, 2), (123456, 77)):
                shares = [mg.candidate_share(frac_local, cell, call, r, world) for r in range(world)]
                assert sum(shares) == int(frac_local * world * world + 0.5)
                assert max(shares) - min(shares) <= 1
