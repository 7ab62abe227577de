"""Multi-GPU plumbing: one process per GPU, parcels (or whole columns) block-partitioned over the ranks with no
halo and no exchange step; the only collective is the all-reduce of the per-slot moment sums (the conservation
diagnostic the reference computes serially in test/examples/utils/netcdf_helpers.jl:34-42)."""
from typing import Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition [lo, hi) of ``n_total`` parcels: the first ``n_total % world`` ranks get one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_columns(n_columns: int, nz: int, rank: int, world: int) -> Tuple[int, int]:
    """Cell range [lo, hi) of whole columns (cells of a column couple through the level above, rainshaft_helpers.jl:83-85,
    so a column is never split → no halo)."""
    lo, hi = shard_range(n_columns, rank, world)
    return lo * nz, hi * nz


def all_reduce_moment_sums(sums):
    """Sum a torch tensor of per-slot moment sums over all ranks, in place (NCCL on GPU tensors, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def total_mass(sums, NProgMoms) -> float:
    """Σ over modes of the first moment (conserved by coalescence)."""
    out, s = 0.0, 0
    for n in NProgMoms:
        out += float(sums[s + 1])
        s += n
    return out
