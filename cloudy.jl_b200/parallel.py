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


def init_comm(ctx, rank: int = None, world: int = None, group=None):
    """Create the context's NCCL communicator (cloudy_comm_init) with torch.distributed as the bootstrap: rank 0 draws the
    unique id (cloudy_comm_unique_id) and broadcasts its 128 bytes through the already initialised process group (gloo or
    nccl).  A Julia host would ship the same bytes with MPI.jl or a shared file."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from . import _lib as L
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    buf = (C.c_char * 128)()
    if world > 1:
        if rank == 0:
            L.check(L.load().cloudy_comm_unique_id(C.cast(buf, C.c_void_p)))
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda" if on_gpu else "cpu")
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        C.memmove(buf, raw, 128)
    L.check(L.load().cloudy_comm_init(ctx.handle, int(world), int(rank), C.cast(buf, C.c_void_p)))
    return rank, world
