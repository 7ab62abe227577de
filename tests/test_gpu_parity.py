"""CUDA path vs the oracle, through the C ABI (ctypes).  Tolerance: tendencies rtol 1e-9 (BASELINE.json
north_star), applied as |Δ| <= rtol*max(|ref|, Σ|terms|)."""
import math

import numpy as np
import pytest

from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params, tendency_close

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def cb():
    import cloudy_b200
    return cloudy_b200


def _check_box(cb, par, state, n_check, lanes=(8,)):
    opar = oracle_params(par)
    model = cb.CoalescenceModel(par)
    ref = np.zeros((n_check, state.shape[1]))
    sc = np.zeros_like(ref)
    for i in range(n_check):
        ref[i], sc[i] = O.rhs_coal(state[i], opar, return_scale=True)
    for ln in lanes:
        model.ctx.set_lanes(ln)
        got = model.coal_tendency_host(state)
        ok, worst = tendency_close(got[:n_check], ref, sc, RTOL)
        assert ok, f"lanes={ln}: worst scaled error {worst:.3e}"
    model.ctx.set_lanes(0)
    return got


def test_c1_smoluchowski(cb):
    from cloudy_b200 import workloads as W
    par, state = W.c1_smoluchowski()
    _check_box(cb, par, state, 1)


def test_c2_gamma_exp(cb):
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_exp(n_parcels=3000)
    _check_box(cb, par, state, 400, lanes=(4, 8, 16, 32))


def test_c2_gamma_gamma(cb):
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_gamma(n_parcels=700)
    got = _check_box(cb, par, state, 200)
    kat = (-5623499.479946031, -3.975692677217648e-4, -6.433699758256767e-14, 623494.4299459805,
           3.975692677217648e-4, 2.6435719760256764e-13)  # SURVEY Appendix B KAT-D
    assert np.allclose(got[0], kat, rtol=1e-10, atol=0)
