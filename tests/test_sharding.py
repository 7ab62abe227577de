"""The N>1 path on CPU: world_size-2 gloo processes shard an ensemble, reduce the moment sums and agree with the
single-process result.  (The device kernels are exercised by the -m gpu tests; this covers the host logic.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from cloudy_b200 import parallel, workloads as W
    par, state = W.c2_gamma_exp(n_parcels=n_total)
    lo, hi = parallel.shard_range(n_total, rank, world)
    local = torch.from_numpy(state[lo:hi].sum(axis=0).copy())
    parallel.all_reduce_moment_sums(local)
    mass = parallel.total_mass(local, par.NProgMoms)
    q.put((rank, lo, hi, local.numpy().copy(), mass))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_allreduce():
    from cloudy_b200 import workloads as W
    n_total = 10007
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    par, state = W.c2_gamma_exp(n_parcels=n_total)
    want = state.sum(axis=0)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n_total  # contiguous cover, no overlap
    for r in res:
        assert np.allclose(r[3], want, rtol=1e-12)
        assert abs(r[4] - (want[1] + want[4])) <= 1e-12 * abs(want[1] + want[4])


def test_shard_helpers():
    from cloudy_b200 import parallel
    for n, w in ((10, 3), (64, 8), (7, 8), (0, 2)):
        spans = [parallel.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    lo, hi = parallel.shard_columns(4096, 256, 3, 8)
    assert lo % 256 == 0 and hi % 256 == 0 and hi - lo == 512 * 256
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)
