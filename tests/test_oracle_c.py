"""The C/OpenMP restatement (cpu_baseline) against the scipy oracle."""
import numpy as np

from oracle import c_oracle, cloudy_oracle as O
from tests.oracle_bridge import oracle_params, tendency_close


def _cfg(par):
    import cloudy_b200 as cb
    kinds = tuple(d.kind for d in par.pdists)
    return cb.build_config(kinds, par.coal_data, norms=par.norms)


def test_c_oracle_matches_scipy_oracle_c2():
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_exp(n_parcels=64)
    got = c_oracle.rhs_coal_batch(_cfg(par), state, n_threads=2)
    opar = oracle_params(par)
    for i in range(24):
        ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
        ok, worst = tendency_close(got[i], ref, sc, 1e-11)
        assert ok, worst


def test_c_oracle_matches_scipy_oracle_mono_and_long():
    from cloudy_b200 import workloads as W
    for gen in (W.mono_gamma, W.long_kernel_two_modes):
        par, state = gen(n_parcels=16)
        got = c_oracle.rhs_coal_batch(_cfg(par), state, n_threads=1)
        opar = oracle_params(par)
        for i in range(8):
            ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
            ok, worst = tendency_close(got[i], ref, sc, 1e-11)
            assert ok, (gen.__name__, worst)
