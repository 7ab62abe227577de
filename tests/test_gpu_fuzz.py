"""Randomised configurations: every thread-per-parcel kernel instance (and a few shapes without one) against the oracle,
and the two independent CUDA implementations (thread-per-parcel vs lane-cooperative) against each other."""
import numpy as np
import pytest

from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params, tendency_close

pytestmark = pytest.mark.gpu

TPP_SHAPES = [(1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (2, 1), (2, 2), (2, 3), (2, 4), (2, 5), (3, 1), (3, 2), (3, 3), (3, 4), (3, 5),
              (4, 1), (4, 2), (4, 3)]
OTHER_SHAPES = [(4, 4), (4, 5)]  # no thread-per-parcel instance: lane-cooperative kernel only


@pytest.mark.parametrize("N,P", TPP_SHAPES + OTHER_SHAPES)
def test_random_configurations(N, P):
    import cloudy_b200 as cb
    from cloudy_b200 import workloads as W
    rng = np.random.default_rng(1000 * N + P)
    for trial in range(3):
        par, state = W.random_model(rng, N, P, n_parcels=96)
        opar = oracle_params(par)
        model = cb.CoalescenceModel(par)
        has_tpp = (N, P) in TPP_SHAPES
        outs = {}
        for lanes in ((1, 8, 32) if has_tpp else (0, 4)):
            model.ctx.set_lanes(lanes)
            outs[lanes] = model.coal_tendency_host(state)
        model.ctx.set_lanes(0)
        n_check = 10
        for i in range(n_check):
            ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
            for lanes, got in outs.items():
                ok, worst = tendency_close(got[i], ref, sc, 1e-9)
                assert ok, (N, P, trial, [d.kind for d in par.pdists], par.coal_data.dist_thresholds, lanes, i, worst)
        # the two CUDA implementations agree on the whole batch (cancellation-aware: compare against the moment scale)
        keys = list(outs)
        a, b = outs[keys[0]], outs[keys[-1]]
        scale = np.abs(a).max(axis=1, keepdims=True) + 1e-300
        assert np.all(np.abs(a - b) <= 1e-9 * scale + 1e-7 * np.abs(a))


@pytest.mark.parametrize("N,P", [sh for sh in TPP_SHAPES if sh[0] >= 2])
def test_random_moving_threshold_configurations(N, P):
    """MovingThreshold (Coalescence.jl:152-185) on every thread-per-parcel shape: random percentiles per mode, thresholds
    on both sides of 1 (unit grid and per-parcel grid), against the oracle"""
    import cloudy_b200 as cb
    from cloudy_b200 import workloads as W
    rng = np.random.default_rng(7000 + 10 * N + P)
    for trial in range(2):
        par, state = W.random_model(rng, N, P, n_parcels=96, moving=True)
        opar = oracle_params(par)
        got = cb.CoalescenceModel(par).coal_tendency_host(state)
        for i in range(8):
            ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
            ok, worst = tendency_close(got[i], ref, sc, 1e-9)
            assert ok, (N, P, trial, [d.kind for d in par.pdists], par.coal_data.dist_thresholds, i, worst)
