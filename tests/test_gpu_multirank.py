"""The conservation all-reduce behind the C ABI (cloudy_comm_init / cloudy_moment_sums_allreduce), one process per GPU.
Needs two GPUs for the multi-rank case (skipped on a one-GPU box); the one-rank communicator runs anywhere."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_rank_communicator_returns_local_sums():
    import cloudy_b200 as cb
    from cloudy_b200 import workloads as W
    from cloudy_b200.parallel import init_comm
    par, state = W.c2_gamma_exp(n_parcels=5000)
    ctx = cb.Context(0)
    assert init_comm(ctx, rank=0, world=1) == (0, 1)
    model = cb.CoalescenceModel(par, ctx=ctx)
    u = model.ensemble(5000).upload(state)
    got = model.moment_sums_allreduce(u)
    assert np.array_equal(got, model.moment_sums(u))
    model.moment_sums_allreduce(u, wait=False)
    assert np.array_equal(model.moment_sums_fetch(), got)
    with pytest.raises(cb.CloudyError):
        model.moment_sums_fetch()  # nothing pending


@pytest.mark.parametrize("world", [2])
def test_allreduce_across_ranks(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mp_allreduce_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTIRANK_OK" in res.stdout
