"""Worker of tests/test_gpu_multirank.py — launched as
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mp_allreduce_worker.py
One process per GPU.  torch.distributed (gloo) is only the bootstrap that ships the 128-byte NCCL id; the conservation
all-reduce itself runs inside libcloudy_b200.so (cloudy_comm_init / cloudy_moment_sums_allreduce)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    import cloudy_b200 as cb
    from cloudy_b200 import workloads as W
    from cloudy_b200.parallel import init_comm, shard_range, total_mass

    ctx = cb.Context(local)
    assert init_comm(ctx) == (rank, world)
    n_total = 40000 * world + 17            # ragged shards
    par, state = W.c2_gamma_exp(n_parcels=n_total)
    lo, hi = shard_range(n_total, rank, world)
    model = cb.CoalescenceModel(par, ctx=ctx)
    model.ctx.set_regime_sort(True)
    u = model.ensemble(hi - lo).upload(state[lo:hi])
    sums0 = model.moment_sums_allreduce(u)
    host = state.sum(axis=0)
    assert np.allclose(sums0, host, rtol=1e-12, atol=0), (sums0, host)
    # every rank holds the same global sums, bit for bit
    gathered = [None] * world
    dist.all_gather_object(gathered, sums0.tobytes())
    assert all(g == gathered[0] for g in gathered)
    # 10 fused steps: coalescence conserves the total mass (sum over modes of M1), number decreases
    model.ssprk33_steps(u, 0.01, 10, cb.MODEL_BOX)
    model.moment_sums_allreduce(u, wait=False)     # enqueue only ...
    du = model.ensemble(hi - lo)
    model.coal_tendency(u, du)                     # ... the next evaluation overlaps the collective
    sums1 = model.moment_sums_fetch()
    m0, m1 = total_mass(sums0, par.NProgMoms), total_mass(sums1, par.NProgMoms)
    assert abs(m1 - m0) <= 1e-11 * abs(m0), (m0, m1)
    assert sums1[0] + sums1[3] < sums0[0] + sums0[3]
    # the same run on the whole ensemble by one rank gives the same global sums
    if rank == 0:
        ctx1 = cb.Context(local)
        model1 = cb.CoalescenceModel(par, ctx=ctx1)
        v = model1.ensemble(n_total).upload(state)
        model1.ssprk33_steps(v, 0.01, 10, cb.MODEL_BOX)
        single = model1.moment_sums(v)
        assert np.allclose(sums1, single, rtol=1e-12, atol=0), (sums1, single)
    dist.barrier()
    L = cb._lib
    import ctypes as C
    ver = C.c_int32()
    L.check(L.load().cloudy_comm_info(ctx.handle, None, None, C.byref(ver)))
    if rank == 0:
        print(f"MULTIRANK_OK world={world} nccl={ver.value} mass={m0!r}->{m1!r}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
