"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
 - CPU: the oracle still reproduces them (guards against oracle drift);
 - GPU: the CUDA path reproduces them through the C ABI at the parity tolerance."""
import glob
import os

import numpy as np
import pytest

from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params, tendency_close

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    from tests.golden.make_golden import CASES
    return CASES


def test_fixtures_exist():
    names = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "*.npz"))}
    assert set(_cases()) | {"c3_rainshaft_column"} <= names


@pytest.mark.parametrize("name", ["c1_smoluchowski", "c2_gamma_exp", "mono_gamma", "moving_gamma_exp"])
def test_oracle_reproduces_fixture(name):
    gen, kw, n = _cases()[name]
    par, state = gen(**kw)
    fx = np.load(os.path.join(HERE, name + ".npz"))
    assert np.array_equal(fx["state"], state[:n])
    opar = oracle_params(par)
    for i in range(0, n, max(1, n // 8)):
        ref = O.rhs_coal(state[i], opar)
        assert np.allclose(ref, fx["tendency"][i], rtol=1e-12, atol=1e-12 * fx["scale"][i].max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(["c1_smoluchowski", "c2_gamma_exp", "c2_gamma_gamma", "mono_gamma", "long_kernel_two_modes",
                                         "c4_three_modes", "moving_four_modes", "moving_gamma_exp"]))
def test_cuda_reproduces_fixture(name):
    import cloudy_b200 as cb
    gen, kw, n = _cases()[name]
    par, _ = gen(**kw)
    fx = np.load(os.path.join(HERE, name + ".npz"))
    model = cb.CoalescenceModel(par)
    got = model.coal_tendency_host(fx["state"])
    ok, worst = tendency_close(got, fx["tendency"], fx["scale"], 1e-9)
    assert ok, worst


@pytest.mark.gpu
def test_cuda_reproduces_rainshaft_fixture():
    import cloudy_b200 as cb
    from cloudy_b200 import workloads as W
    par, _ = W.c3_rainshaft(n_columns=1, nz=20)
    fx = np.load(os.path.join(HERE, "c3_rainshaft_column.npz"))
    rhs = cb.make_rainshaft_rhs(cb.AnalyticalCoalStyle())
    m = fx["state"].copy()
    got = rhs(m, par, 0.0)
    assert np.array_equal(m, fx["clipped"])
    ok, worst = tendency_close(got, fx["rhs"], fx["scale"], 1e-9)
    assert ok, worst
    model = cb.CoalescenceModel(par, nz=20)
    u = model.ensemble(20).upload(fx["state"])
    model.ssprk33_steps(u, par.dt, 10, cb.MODEL_RAINSHAFT)
    after = u.download()
    ref = fx["after10"]
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=0, keepdims=True) * 1e-6)
    assert np.all(np.abs(after - ref) <= 1e-7 * scale)
