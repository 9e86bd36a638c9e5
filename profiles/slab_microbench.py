#!/usr/bin/env python
"""Half-sweep cost of the slab-decomposed SOR (run under torchrun, one rank per GPU):
   torchrun --nproc-per-node G profiles/slab_microbench.py [ni nj nk iters]
Prints, on rank 0, microseconds per colour half-sweep for the slab solve and for a replicated solve of one slab's size."""
import importlib, os, sys, time
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pg = importlib.import_module(bench.PKG + ".picgpu")
pg.init(local)
ni, nj, nk, iters = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (256, 256, 256, 200)))


def all_gather_bytes(b):
    t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.cpu().numpy().tobytes() for o in out]


def run(n_i, slab):
    w = pg.World(n_i, nj, nk, (0, 0, 0), (1e-4 * (n_i - 1), 1e-4 * (nj - 1), 1e-4 * (nk - 1)))
    w.setTime(1e-12, 1 << 30); w.computeObjectID()
    sol = pg.PotentialSolver(w, 100000, 1e-30); sol.setReferenceValues(0, 0, 1e20)
    if slab:
        sol.enableSlabs(rank, world, all_gather_bytes)
    sol.iterate(iters); pg.synchronize()                      # warm-up with the same batch size (captures the CUDA graph)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter(); sol.iterate(iters); pg.synchronize(); dt = time.perf_counter() - t0
    sol.close(); w.close()
    return dt / (2 * iters) * 1e6


res = {}
if world > 1:
    res["slab %dx%dx%d over %d ranks" % (ni, nj, nk, world)] = run(ni, True)
res["replicated %dx%dx%d (one slab's planes)" % (ni // world, nj, nk)] = run(max(ni // world, 4), False)
if rank == 0:
    for k, v in res.items():
        print("%-50s %8.2f us per half-sweep" % (k, v))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
