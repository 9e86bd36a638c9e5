"""Multi-GPU plumbing of the particle loop (one process per GPU, torch.distributed).

Partitioning: particles are split evenly by index across the G ranks; every rank keeps the full grid and
deposits its particles into a full-grid int64 fixed-point accumulator; the accumulators are summed with one all-reduce
per species (NCCL over NVLink on the GPU box, gloo in the CPU tests) and only then turned into densities.  Because the
accumulators are integers, the reduced grid is bit-identical to a single-GPU deposit of all particles, for any G and any
reduction order.  Poisson is solved on slabs (csrc/poisson.cu) or redundantly; Monte-Carlo collisions stay per rank, a cell's
candidates are dealt out to the ranks (candidate_share).

This module holds the host-side rules that must agree on every rank; it contains no compute.
"""
import math


def split_count(n_total, rank, world):
    """Particles owned by `rank` when n_total are split evenly by index (csrc/source.cu uses the same rule)."""
    per, rem = divmod(int(n_total), int(world))
    return per + (1 if rank < rem else 0)


def common_scale(S_local, world, reduce_min):
    """The fixed-point scale every rank must use: the most conservative local calibration, lowered by ceil(log2 G) so that
    the sum over G ranks keeps the same headroom.  reduce_min(int) -> int is the all-reduce(MIN) of the caller's backend."""
    return int(reduce_min(int(S_local))) - int(math.ceil(math.log2(world))) if world > 1 else int(S_local)


def candidate_share(frac_local, cell, call, rank, world):
    """Monte-Carlo candidates rank `rank` tries in `cell` (csrc/mcc.cu, csrc/dsmc.cu use the same rule).  Cells are not owned by a
    GPU: the cell's candidate count is estimated from the LOCAL populations (frac_local is the reference's
    expression on them, bilinear in the counts, hence x G^2), rounded once with the reference's int(x + 0.5)
    (Interactions.cpp:646-647) and dealt out: n // G to every rank, the n % G left over to a subset that rotates with the cell
    and the call number.  Rounding each rank's share instead would drop every cell whose share is below one half."""
    n_tot = int(frac_local * world * world + 0.5)
    return n_tot // world + (1 if (cell + call + rank) % world < n_tot % world else 0)


def common_ceiling(value, reduce_max):
    """W_sigma_v_rel_max / sigma_v_rel_max must be the same on every rank (it normalises the acceptance probability): the
    all-reduce(MAX) of the per-rank values after every apply.  reduce_max(float) -> float."""
    return float(reduce_max(float(value)))


class CudaArray:
    """Zero-copy view of a device buffer owned by libpicgpu.so for torch.as_tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def fixed_view(torch, species, field_id):
    """torch int64 tensor aliasing the species' raw fixed-point accumulator (all-reduced in place)."""
    ptr, nbytes = species.device_ptr(field_id)
    return torch.as_tensor(CudaArray(ptr, nbytes // 8, "<i8"), device="cuda")
