"""Regime order of resident ensembles (cloudy_state_regime_sort, regime_sort.cuh): the parcels are moved, the results are not.
Everything here is bit-exact: a parcel's tendency does not depend on its position or its warp-mates."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import cloudy_b200
    return cloudy_b200


def _plain_tendency(cb, par, state):
    model = cb.CoalescenceModel(par)
    model.ctx.set_regime_sort(False)
    n = state.shape[0]
    u = model.ensemble(n).upload(state); du = model.ensemble(n)
    model.coal_tendency(u, du)
    assert u.order() is None and du.order() is None
    return model, du.download()


@pytest.mark.parametrize("n", [4096, 4097, 12289, 70000])
def test_sort_roundtrip_and_order(cb, n):
    """sorting moves whole parcels: the order is a permutation, the download restores the host order bit for bit, the
    raw device buffer holds parcel order[i] at position i, and sorting twice composes the orders (stable sort: the
    second sort of unchanged data is the identity on positions)"""
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_exp(n_parcels=n)
    model = cb.CoalescenceModel(par)
    u = model.ensemble(n).upload(state)
    assert u.order() is None
    u.regime_sort()
    order = u.order()
    assert order is not None and np.array_equal(np.sort(order), np.arange(n))
    assert not np.array_equal(order, np.arange(n))
    assert np.array_equal(u.download(), state)
    # raw SoA buffer: position i holds parcel order[i]
    raw = np.empty((u.n_slots, u.stride))
    from cloudy_b200.ensemble import _DeviceBuffer
    model.ctx.sync()
    _DeviceBuffer.memcpy_d2h(raw, u.device_ptr(), raw.nbytes)
    assert np.array_equal(raw[:, :n].T, state[order])
    u.regime_sort()
    assert np.array_equal(u.order(), order)
    assert np.array_equal(u.download(), state)
    u.upload(state)
    assert u.order() is None


def test_tendency_on_resident_order_is_bit_identical(cb):
    """cloudy_coal_tendency sorts an unsorted input once (regime sort on), the output inherits the order, downloads agree
    bit for bit with the unsorted evaluation; the host-buffer pipeline (permutation sort) agrees too"""
    from cloudy_b200 import workloads as W
    for gen, n in ((W.c2_gamma_exp, 50000), (W.c4_three_modes, 9000), (W.moving_gamma_exp, 20000)):
        par, state = gen(n_parcels=n)
        model, ref = _plain_tendency(cb, par, state)
        model.ctx.set_regime_sort(True)
        s0 = model.ctx.sort_count()
        u = model.ensemble(n).upload(state); du = model.ensemble(n)
        model.coal_tendency(u, du)
        assert model.ctx.sort_count() == s0 + 1
        assert u.order() is not None and np.array_equal(u.order(), du.order())
        got = du.download()
        model.coal_tendency(u, du)            # already resident: no second sort
        assert model.ctx.sort_count() == s0 + 1
        assert np.array_equal(got, ref) and np.array_equal(du.download(), ref)
        assert np.array_equal(u.download(), state)
        assert np.array_equal(model.coal_tendency_host(state), ref)
        model.ctx.set_regime_sort(False)


def test_fused_steps_with_resort_interval(cb):
    """the stepper refreshes the order every `resort_interval` steps; integrated states are bit-identical whatever the interval"""
    from cloudy_b200 import workloads as W
    n = 30000
    par, state = W.c2_gamma_exp(n_parcels=n)
    model = cb.CoalescenceModel(par)
    model.ctx.set_regime_sort(False)
    u0 = model.ensemble(n).upload(state)
    model.ssprk33_steps(u0, 0.05, 5, cb.MODEL_BOX)
    plain = u0.download()
    model.ctx.set_regime_sort(True)
    for interval, want_sorts in ((1, 5), (2, 3), (10, 1)):
        model.ctx.set_resort_interval(interval)
        s0 = model.ctx.sort_count()
        u = model.ensemble(n).upload(state)
        model.ssprk33_steps(u, 0.05, 5, cb.MODEL_BOX)
        assert model.ctx.sort_count() - s0 == want_sorts
        assert np.array_equal(u.download(), plain, equal_nan=True)
        # split into two calls: the age of the order carries over
        s0 = model.ctx.sort_count()
        u = model.ensemble(n).upload(state)
        model.ssprk33_steps(u, 0.05, 2, cb.MODEL_BOX)
        model.ssprk33_steps(u, 0.05, 3, cb.MODEL_BOX)
        assert model.ctx.sort_count() - s0 == want_sorts
        assert np.array_equal(u.download(), plain, equal_nan=True)
    model.ctx.set_resort_interval(10)
    model.ctx.set_regime_sort(False)


def test_sums_diagnostics_and_copy_follow_the_order(cb):
    from cloudy_b200 import workloads as W
    n = 20000
    par, state = W.c2_gamma_exp(n_parcels=n)
    model = cb.CoalescenceModel(par)
    model.ctx.set_regime_sort(False)
    u = model.ensemble(n).upload(state)
    nq_plain = model.standard_N_q(u, 0.3, normalized=True)
    du = model.ensemble(n)
    model.cond_evap(u, du, 0.01, 1e-2)
    ce_plain = du.download()
    u.regime_sort()
    # per-parcel diagnostics come back indexed by parcel, not by position
    assert np.array_equal(model.standard_N_q(u, 0.3, normalized=True), nq_plain)
    model.cond_evap(u, du, 0.01, 1e-2)
    assert np.array_equal(du.order(), u.order())
    assert np.array_equal(du.download(), ce_plain)
    # the sort is deterministic: same order, same (order-dependent) rounding of the sums, run after run
    sums = model.moment_sums(u)
    u2 = model.ensemble(n).upload(state).regime_sort()
    assert np.array_equal(u2.order(), u.order())
    assert np.array_equal(model.moment_sums(u2), sums)
    assert np.allclose(sums, state.sum(axis=0), rtol=1e-12)
    # a copy holds the same parcels at the same positions
    v = model.ensemble(n)
    from cloudy_b200 import _lib as L
    L.check(L.load().cloudy_state_copy(model.ctx.handle, u.handle, v.handle))
    assert np.array_equal(v.order(), u.order()) and np.array_equal(v.download(), state)


def test_column_states_keep_their_order(cb):
    from cloudy_b200 import workloads as W
    from cloudy_b200._lib import CloudyError
    par, cols = W.c3_rainshaft(64, 64)
    model = cb.CoalescenceModel(par, nz=64)
    u = model.ensemble(64 * 64).upload(cols.reshape(-1, 6))
    with pytest.raises(CloudyError):
        u.regime_sort()


def test_context_reconfigured_with_another_slot_count(cb):
    """one context, two configurations with different numbers of moments and the same ensemble size: the sort scratch of
    the first (5 slots) must not be reused for the second (12 slots) — results equal those of a fresh context"""
    from cloudy_b200 import workloads as W
    n = 270000  # above the automatic regime-sort threshold
    ctx = cb.Context(0)
    par1, st1 = W.c2_gamma_exp(n_parcels=n)
    m1 = cb.CoalescenceModel(par1, ctx=ctx)
    u1 = m1.ensemble(n).upload(st1); d1 = m1.ensemble(n)
    m1.coal_tendency(u1, d1)
    assert u1.order() is not None
    got1 = d1.download()
    par2, st2 = W.moving_four_modes(n_parcels=n)
    m2 = cb.CoalescenceModel(par2, ctx=ctx)
    u2 = m2.ensemble(n).upload(st2); d2 = m2.ensemble(n)
    m2.coal_tendency(u2, d2)
    got2 = d2.download()
    fresh = cb.CoalescenceModel(par2, ctx=cb.Context(0))
    uf = fresh.ensemble(n).upload(st2); df = fresh.ensemble(n)
    fresh.coal_tendency(uf, df)
    assert np.array_equal(got2, df.download())
    fresh1 = cb.CoalescenceModel(par1, ctx=cb.Context(0))
    assert np.array_equal(got1, fresh1.coal_tendency_host(st1))
