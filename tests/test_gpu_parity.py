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


def _check_box(cb, par, state, n_check, lanes=(0,)):
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
    _check_box(cb, par, state, 400, lanes=(1, 4, 8, 16, 32))


def test_c2_gamma_gamma(cb):
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_gamma(n_parcels=700)
    got = _check_box(cb, par, state, 200)
    kat = (-5623499.479946031, -3.975692677217648e-4, -6.433699758256767e-14, 623494.4299459805,
           3.975692677217648e-4, 2.6435719760256764e-13)  # SURVEY Appendix B KAT-D
    assert np.allclose(got[0], kat, rtol=1e-10, atol=0)


def test_mono_gamma_and_long_kernel(cb):
    from cloudy_b200 import workloads as W
    for gen in (W.mono_gamma, W.long_kernel_two_modes):
        par, state = gen(n_parcels=300)
        _check_box(cb, par, state, 120, lanes=(1, 4, 8))


def test_c4_three_gamma_modes_order4(cb):
    from cloudy_b200 import workloads as W
    par, state = W.c4_three_modes(n_parcels=1000)
    _check_box(cb, par, state, 60, lanes=(1, 4, 8, 32))


def test_ragged_and_empty_batches(cb):
    """sizes that are not a multiple of the tile, a single parcel, and the empty batch"""
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_exp(n_parcels=67)
    full = _check_box(cb, par, state, 67)
    model = cb.CoalescenceModel(par)
    for lanes in (1, 8):
        model.ctx.set_lanes(lanes)
        full = model.coal_tendency_host(state)
        for n in (1, 31, 33):
            got = model.coal_tendency_host(state[:n])
            assert np.array_equal(got, full[:n])  # a parcel's result does not depend on its neighbours: bit-identical
    model.ctx.set_lanes(0)
    assert model.coal_tendency_host(np.zeros((0, 5))).shape == (0, 5)


def test_all_empty_and_extreme_parcels(cb):
    """empty modes (n = 0 fallback), huge / tiny scale parameters, shape parameter at its clamps"""
    from cloudy_b200 import workloads as W
    par, _ = W.c2_gamma_exp(n_parcels=8)
    nf = np.array([1e6, 1e-3, 1e-12, 1e6, 1e-3])
    rows = [
        [0, 0, 0, 0, 0],                       # everything empty
        [100, 10, 2, 0, 0],                    # rain empty
        [0, 0, 0, 1, 1],                       # cloud empty
        [100, 10, 1.0000001, 1, 1],            # k clamps at 10 (variance ~ 0)
        [100, 10, 1e6, 1, 1],                  # k tiny
        [100, 1e-3, 2e-8, 1, 1],               # θ = 1e-5: threshold far in the tail (z up to 5e4)
        [100, 1e4, 2e6, 1, 1],                 # θ = 100: threshold far below the mean
        [1e-3, 1e-4, 2e-5, 1e-9, 1e-7],        # small numbers
        [100, 60, 40, 1, 1],                   # X = x_th/θ around the series/continued-fraction switch
        [100, 2.5, 0.07, 1, 1],
    ]
    state = np.array(rows, dtype=np.float64) * nf
    _check_box(cb, par, state, len(rows), lanes=(1, 4, 8, 16, 32))


def test_full_size_properties(cb):
    """BASELINE configs[1] at full size (1,048,576 parcels): size-independent properties — total mass tendency
    cancels, number tendency is negative, permutation invariance, determinism."""
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_exp()
    model = cb.CoalescenceModel(par)
    got = model.coal_tendency_host(state)
    mass = got[:, 1] + got[:, 4]
    scale = np.abs(got[:, 1]) + np.abs(got[:, 4]) + 1e-300
    assert np.all(np.abs(mass) <= 1e-9 * scale + 1e-9 * np.abs(state[:, 1] + state[:, 4]))
    assert np.all(got[:, 0] + got[:, 3] <= 0)
    assert np.all(np.isfinite(got))
    perm = np.random.default_rng(1).permutation(state.shape[0])
    got_p = model.coal_tendency_host(state[perm])
    assert np.array_equal(got_p, got[perm])
    assert np.array_equal(model.coal_tendency_host(state), got)
    # spot-check against the oracle across the ensemble
    opar = oracle_params(par)
    for i in range(0, state.shape[0], 16384):
        ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
        ok, worst = tendency_close(got[i], ref, sc, RTOL)
        assert ok, (i, worst)


def test_sedimentation_flux_batched(cb):
    from cloudy_b200 import workloads as W
    par, cols = W.c3_rainshaft(n_columns=3, nz=16)
    rng = np.random.default_rng(3)
    state = cols.reshape(-1, 6).copy()
    state[:, 3:] = state[:, :3] * rng.uniform(0.0, 1e-3, (state.shape[0], 1)) * np.array([1e-2, 1.0, 50.0])
    model = cb.CoalescenceModel(par, nz=16)
    u = model.ensemble(state.shape[0]).upload(state)
    fl = model.ensemble(state.shape[0])
    model.sedimentation_flux(u, fl)
    got = fl.download()
    opar = oracle_params(par)
    for i in range(state.shape[0]):
        ref = O.sedimentation_flux_state(state[i], opar)
        assert np.allclose(got[i], ref, rtol=1e-12, atol=0), (i, got[i], ref)


def _rain_state(cols, seed=5):
    """give the rain mode a consistent Gamma content and sprinkle negatives (the RHS clips them in place)"""
    rng = np.random.default_rng(seed)
    st = cols.copy()
    shp = st.shape[:-1]
    frac = rng.uniform(0.0, 2e-3, shp) * (st[..., 1] > 0)          # share of the cloud mass that is rain
    th = np.exp(rng.uniform(np.log(1.0), np.log(8.0), shp)) * 1e-9  # kg
    k = rng.uniform(0.8, 3.0, shp)
    m1 = st[..., 1] * frac
    st[..., 3] = m1 / (th * k)
    st[..., 4] = m1
    st[..., 5] = m1 * th * (k + 1)
    neg = rng.random(st.shape) < 0.02
    st[neg] = -np.abs(st[neg]) * 1e-3 - 1e-30
    return st


def test_c3_rainshaft_rhs(cb):
    from cloudy_b200 import workloads as W
    for nz, ncol in ((20, 3), (37, 2), (256, 1)):
        par, cols = W.c3_rainshaft(n_columns=ncol, nz=nz)
        st = _rain_state(cols)
        rhs = cb.make_rainshaft_rhs(cb.AnalyticalCoalStyle())
        m = st.copy()
        got = rhs(m, par, 0.0)
        opar = oracle_params(par)
        for c in range(ncol):
            mo = st[c].copy()
            ref, scale = O.rainshaft_rhs(mo, opar, return_scale=True)
            assert np.array_equal(m[c], mo)  # clipped in place, identically
            ok, worst = tendency_close(got[c], ref, scale, RTOL)
            assert ok, (nz, c, worst)
            assert np.abs(ref).max() > 0


def test_ssprk33_box_matches_oracle_run(cb):
    """integrated moments rtol 1e-7 after the full run (north_star).  box_gamma_mixture.jl: 12 steps of dt = 10 from the
    script's own initial condition; random ensemble members (10x denser) with dt = 1."""
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_gamma(n_parcels=40)
    model = cb.CoalescenceModel(par)
    opar = oracle_params(par)
    for dt, idx in ((10.0, [0]), (1.0, list(range(1, 40, 6)))):
        u = model.ensemble(state.shape[0]).upload(state)
        model.ssprk33_steps(u, dt, 12, cb.MODEL_BOX)
        got = u.download()
        for i in idx:
            ref = O.ssprk33(lambda m: O.rhs_coal(m, opar), state[i], dt, 12)
            assert np.all(np.isfinite(ref))
            assert np.allclose(got[i], ref, rtol=1e-7, atol=0), (dt, i, got[i], ref)
    # odd and even step counts leave the result in the caller's ensemble
    u2 = model.ensemble(state.shape[0]).upload(state)
    model.ssprk33_steps(u2, 1.0, 5, cb.MODEL_BOX)
    model.ssprk33_steps(u2, 1.0, 7, cb.MODEL_BOX)
    assert np.allclose(u2.download(), got, rtol=1e-13, atol=0)


def test_ssprk33_box_moving_threshold_matches_oracle_run(cb):
    """box_gamma_mix_moving.jl: 4 Gamma modes, percentile thresholds re-evaluated at every stage (Coalescence.jl:152-185),
    integrated with SSPRK33; the script's own initial condition and a few random ensemble members"""
    from cloudy_b200 import workloads as W
    par, state = W.moving_four_modes(n_parcels=24)
    model = cb.CoalescenceModel(par)
    opar = oracle_params(par)
    nsteps, dt = 8, 1.0
    u = model.ensemble(state.shape[0]).upload(state)
    model.ssprk33_steps(u, dt, nsteps, cb.MODEL_BOX)
    got = u.download()
    for i in (0, 1, 5, 11, 17):
        ref = O.ssprk33(lambda m: O.rhs_coal(m, opar), state[i], dt, nsteps)
        assert np.all(np.isfinite(ref))
        # rtol 1e-7 on every moment that carries mass; moments many orders below the same-order moment of the dominant
        # mode (freshly seeded modes) are limited by the cancellation in their tendencies
        atol = 1e-10 * np.abs(ref.reshape(4, 3)).max(axis=0)[None, :].repeat(4, axis=0).reshape(-1)
        assert np.all(np.abs(got[i] - ref) <= 1e-7 * np.abs(ref) + atol), (i, got[i], ref)


def test_ssprk33_rainshaft_matches_oracle_run(cb):
    """rainshaft_gamma_mixture.jl:13-49 (nz = 20, dt = 1) for 40 steps, two columns"""
    from cloudy_b200 import workloads as W
    par, cols = W.c3_rainshaft(n_columns=2, nz=20)
    st = _rain_state(cols, seed=9)
    model = cb.CoalescenceModel(par, nz=20)
    flat = st.reshape(-1, 6)
    u = model.ensemble(flat.shape[0]).upload(flat)
    nsteps = 40
    model.ssprk33_steps(u, par.dt, nsteps, cb.MODEL_RAINSHAFT)
    got = u.download().reshape(st.shape)
    opar = oracle_params(par)
    for c in range(2):
        ref = O.ssprk33(lambda m: O.rainshaft_rhs(m, opar), st[c], par.dt, nsteps)
        ref[ref < 0] = 0  # the reference's FSAL evaluation clips the saved state
        assert np.all(np.isfinite(ref))
        scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=0, keepdims=True) * 1e-6)
        assert np.all(np.abs(got[c] - ref) <= 1e-7 * scale), np.max(np.abs(got[c] - ref) / scale)


def test_moment_sums(cb):
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_exp(n_parcels=100003)
    model = cb.CoalescenceModel(par)
    u = model.ensemble(state.shape[0]).upload(state)
    got = model.moment_sums(u)
    assert np.allclose(got, state.sum(axis=0), rtol=1e-12)
    assert np.array_equal(got, model.moment_sums(u))  # deterministic


def test_moving_threshold(cb):
    """get_coal_ints(..., ::MovingThreshold) — Coalescence.jl:152-185; thresholds from per-parcel percentiles
    (compute_thresholds, ParticleDistributions.jl:721-761), grid per parcel."""
    from cloudy_b200 import workloads as W
    par, state = W.moving_four_modes(n_parcels=200)
    _check_box(cb, par, state, 60)
    par, state = W.moving_gamma_exp(n_parcels=300)
    _check_box(cb, par, state, 100)
    par, state = W.moving_gamma_exp(n_parcels=64, percentile=0.5)
    _check_box(cb, par, state, 40)


def _moving_mixed(cb, n, seed, klo=0.5, khi=5.0, percentile=0.97):
    """Gamma + Exponential ensemble whose per-parcel thresholds straddle 1 (unit-grid and own-grid parcels side by side)"""
    from cloudy_b200 import workloads as W
    from cloudy_b200 import _lib as L
    par, _ = W.moving_gamma_exp(n_parcels=8, percentile=percentile)
    rng = np.random.default_rng(seed)
    n1 = W._logu(rng, 1e1, 1e3, n); th1 = W._logu(rng, 0.02, 2.0, n); k1 = W._logu(rng, klo, khi, n)
    n2 = W._logu(rng, 1e-6, 1e0, n); th2 = W._logu(rng, 1.0, 30.0, n)
    m = np.concatenate([W._moments_from_params(L.GAMMA, n1, th1, k1, 3), W._moments_from_params(L.EXPONENTIAL, n2, th2, None, 2)], axis=1)
    return par, m * W._norm_factors(par.NProgMoms, W.NORMS)


def test_moving_threshold_mixed_grids(cb):
    """x_th <= 1 parcels scale the unit grid, x_th > 1 parcels build their own (ParticleDistributions.jl:579-582); both kinds share
    warps here.  Parity with the oracle, and a parcel's result must not depend on its warp-mates or on the regime sort."""
    par, state = _moving_mixed(cb, 6000, 101)
    got = _check_box(cb, par, state, 150)
    model = cb.CoalescenceModel(par)
    perm = np.random.default_rng(5).permutation(state.shape[0])
    got_perm = model.coal_tendency_host(state[perm])
    assert np.array_equal(got_perm, got[perm])
    one_by_one = np.stack([model.coal_tendency_host(state[i:i + 1])[0] for i in range(40)])
    assert np.array_equal(one_by_one, got[:40])
    n = state.shape[0]
    u = model.ensemble(n).upload(state); du = model.ensemble(n)
    model.ctx.set_regime_sort(True)
    model.coal_tendency(u, du)
    model.ctx.set_regime_sort(False)
    assert np.array_equal(du.download(), got)


def test_moving_threshold_small_shape_and_extreme_percentiles(cb):
    """shape parameters below the inverse-incomplete-gamma table (k < 1/4: general Halley iteration) and percentiles far
    from the default"""
    par, state = _moving_mixed(cb, 256, 102, klo=0.03, khi=0.4)
    _check_box(cb, par, state, 60)
    for pct in (0.999, 0.2, 0.01):
        par, state = _moving_mixed(cb, 256, 103, percentile=pct)
        _check_box(cb, par, state, 40)


def test_rainshaft_has_no_moving_threshold_method(cb):
    """the reference's column RHS calls get_coal_ints(style, pdists, coal_data) only (rainshaft_helpers.jl:70)"""
    from cloudy_b200 import workloads as W
    par, cols = W.c3_rainshaft(2, 8)
    cd = cb.CoalescenceData(W.linear_tensor(5.0), par.NProgMoms, (0.97, 1.0), W.NORMS, cb.MovingThreshold())
    par2 = type(par)(**{**vars(par), "coal_data": cd})
    model = cb.CoalescenceModel(par2, nz=8)
    st = cols.reshape(-1, cols.shape[-1])
    u = model.ensemble(st.shape[0]).upload(st); du = model.ensemble(st.shape[0])
    with pytest.raises(Exception, match="MovingThreshold"):
        model.rainshaft_rhs(u, du)


def test_lognormal_mode_with_threshold(cb):
    """box_lognorm_mixture.jl: the reference nests two adaptive QuadGK calls at rtol sqrt(eps) = 1.5e-8, so parity with
    the (scipy.integrate.quad) restatement is asserted at 1e-7 of the term scale, not 1e-9."""
    from cloudy_b200 import workloads as W
    par, state = W.lognormal_mixture(n_parcels=64)
    opar = oracle_params(par)
    model = cb.CoalescenceModel(par)
    got = model.coal_tendency_host(state)
    for i in range(8):
        ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
        ok, worst = tendency_close(got[i], ref, sc, 1e-7)
        assert ok, (i, worst)
    # mass is conserved whatever the quadrature
    assert np.all(np.abs(got[:, 1] + got[:, 4]) <= 1e-12 * (np.abs(got[:, 1]) + np.abs(got[:, 4]) + 1e-300))
    # the same fixed rule in the oracle (closed-form inner integral, 128-point Gauss-Legendre up to the top carried order
    # P + 1): two independent implementations of one rule agree at 1e-11, the 1e-7 above is the distance of that rule (and of
    # the reference's own adaptive one) from the exact integral
    O.LOGNORMAL_FIXED_RULE = (128, par.coal_data.P + 1)
    try:
        for i in range(24):
            ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
            ok, worst = tendency_close(got[i], ref, sc, 1e-11)
            assert ok, (i, worst)
    finally:
        O.LOGNORMAL_FIXED_RULE = None


def test_lognormal_moment_source_helper_scalar(cb):
    """cloudy_moment_source_helper for a Lognormal mode (ParticleDistributions.jl:614-625): the reference's goldens
    (test_ParticleDistributions_correctness.jl:215-218, rtol 1e-3), the oracle's implementation of the same fixed rule at 1e-11,
    a converged evaluation of the closed form at 1e-6, and the adaptive restatement at 1e-6"""
    from cloudy_b200.distributions import LognormalPrimitiveParticleDistribution as LogN
    d = LogN(1.0, 0.5, 2.0)
    od = O.Lognormal(1.0, 0.5, 2.0)
    for (p1, p2, gold) in ((0.0, 0.0, 2.831e-1), (1.0, 0.0, 1.725e-1), (1.0, 1.0, 8.115e-2)):
        got = cb.moment_source_helper(d, p1, p2, 2.5)
        assert abs(got - gold) <= 1e-3 * gold
        assert abs(got - O.moment_source_helper(od, p1, p2, 2.5)) <= 1e-6 * got
    rng = np.random.default_rng(12)
    for _ in range(40):
        mu, sg = rng.uniform(-1.0, 2.0), rng.uniform(0.25, 0.9)
        T = float(np.exp(rng.uniform(mu - sg, mu + 2.5 * sg)))
        p1, p2 = sorted(rng.uniform(0.0, 3.0, 2))
        n = float(np.exp(rng.uniform(-3, 3)))
        got = cb.moment_source_helper(LogN(n, mu, sg), p1, p2, T)
        same_rule = O.moment_source_helper_lognormal_closed(O.Lognormal(n, mu, sg), p1, p2, T, order=128, top_order=max(p1, p2, 0.0))
        assert abs(got - same_rule) <= 1e-11 * abs(same_rule), (mu, sg, T, p1, p2, got, same_rule)
        converged = O.moment_source_helper_lognormal_closed(O.Lognormal(n, mu, sg), p1, p2, T, order=1200)
        assert abs(got - converged) <= 1e-6 * abs(converged), (mu, sg, T, p1, p2, got, converged)


def test_condensation_batched(cb):
    """rhs_condensation! (box_model_helpers.jl:55-67) over an ensemble, scalar and per-parcel supersaturation"""
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_gamma(n_parcels=300)
    opar = oracle_params(par)
    model = cb.CoalescenceModel(par)
    u = model.ensemble(state.shape[0]).upload(state)
    du = model.ensemble(state.shape[0])
    s, xi = 0.01, 1e-6
    model.cond_evap(u, du, s, xi)
    got = du.download()
    for i in range(0, 300, 7):
        ref = O.rhs_condensation(state[i], opar, s, xi)
        assert np.allclose(got[i], ref, rtol=1e-12, atol=0), (i, got[i], ref)


def test_standard_N_q_batched(cb):
    """cloud/rain split of every cell as rainshaft_output computes it (netcdf_helpers.jl:106-121: raw moments, cutoff 5.236e-10)"""
    from cloudy_b200 import workloads as W
    par, cols = W.c3_rainshaft(n_columns=2, nz=20)
    st = _rain_state(cols, seed=11).reshape(-1, 6)
    st[st < 0] = 0
    model = cb.CoalescenceModel(par, nz=20)
    u = model.ensemble(st.shape[0]).upload(st)
    got = model.standard_N_q(u, 5.236e-10, normalized=False)
    for i in range(st.shape[0]):
        pd = [O.update_dist_from_moments(O.Dist(O.GAMMA, 0.0, 1.0, 1.0), st[i, 3 * j:3 * j + 3]) for j in range(2)]
        ref = np.array(O.get_standard_N_q(pd, 5.236e-10))
        scale = max(st[i, 0] + st[i, 3], 1e-300), max(st[i, 0] + st[i, 3], 1e-300), max(st[i, 1] + st[i, 4], 1e-300), max(st[i, 1] + st[i, 4], 1e-300)
        assert np.all(np.abs(got[:, i] - ref) <= 1e-12 * np.array(scale)), (i, got[:, i], ref)


@pytest.mark.parametrize("tile", [None, "4096"])
def test_regime_sort_is_bit_identical(cb, tile, monkeypatch):
    """the optional regime sort only reorders the work: tendencies and fused steps are bit-identical, whether the order is
    built over the whole ensemble or tile by tile (CLOUDY_SORT_TILE; several tiles plus a partial last one here)"""
    from cloudy_b200 import workloads as W
    if tile is not None:
        monkeypatch.setenv("CLOUDY_SORT_TILE", tile)
    else:
        monkeypatch.delenv("CLOUDY_SORT_TILE", raising=False)
    for gen, n, dt in ((W.c2_gamma_exp, 20000, 0.05), (W.c4_three_modes, 6000, 1e-5)):
        par, state = gen(n_parcels=n)
        model = cb.CoalescenceModel(par)
        u = model.ensemble(n).upload(state)
        du = model.ensemble(n)
        model.coal_tendency(u, du)
        ref = du.download()
        model.ctx.set_regime_sort(True)
        model.coal_tendency(u, du)
        got = du.download()
        model.ssprk33_steps(u, dt, 2, cb.MODEL_BOX)
        stepped = u.download()
        model.ctx.set_regime_sort(False)
        u2 = model.ensemble(n).upload(state)
        model.ssprk33_steps(u2, dt, 2, cb.MODEL_BOX)
        assert np.array_equal(got, ref)
        plain = u2.download()
        assert np.isfinite(plain).mean() > 0.99  # extreme synthetic parcels may blow up in both orders alike
        assert np.array_equal(stepped, plain, equal_nan=True)


def test_shape_without_tpp_instance_uses_generic_kernel(cb):
    """(N, P) = (4, 4) has no thread-per-parcel instance: auto mode falls back to the lane-cooperative kernel,
    lanes = 1 (thread per parcel required) reports CLOUDY_ERR_UNSUPPORTED."""
    from cloudy_b200 import workloads as W
    par, state = W.three_modes_order2(n_parcels=200)  # (3, 3): both implementations
    _check_box(cb, par, state, 60, lanes=(0, 1, 4, 8, 16, 32))
    par, state = W.random_model(np.random.default_rng(44), 4, 4, kinds=[cb.GAMMA, cb.EXPONENTIAL, cb.GAMMA, cb.GAMMA], n_parcels=120)
    _check_box(cb, par, state, 30, lanes=(0, 4, 8, 16, 32))
    model = cb.CoalescenceModel(par)
    model.ctx.set_lanes(1)
    with pytest.raises(cb.CloudyError):
        model.coal_tendency_host(state)
    model.ctx.set_lanes(0)


def test_c4_full_size_properties(cb):
    """BASELINE configs[3] at full size (16,777,216 parcels, 3 modes, order-4 tensor): size-independent properties —
    per-parcel mass tendency cancels, the ensemble sums from the device reduction agree with the host, spot checks
    against the oracle across the ensemble."""
    from cloudy_b200 import workloads as W
    n = 1 << 24
    par, state = W.c4_three_modes(n_parcels=n)
    model = cb.CoalescenceModel(par)
    u = model.ensemble(n).upload(state)
    du = model.ensemble(n)
    model.coal_tendency(u, du)
    got = du.download()
    assert np.isfinite(got).all()
    # total mass tendency cancels.  The per-mode mass tendencies are themselves residuals of much larger Q/R/S terms
    # (order-4 tensor, moments up to order 6), so their sum is compared with a LOOSE bound on every parcel and with the
    # tight 1e-9 bound on the bulk; the oracle spot checks below use the exact term scale.
    mass = got[:, 1] + got[:, 4] + got[:, 7]
    scale = np.abs(got[:, 1]) + np.abs(got[:, 4]) + np.abs(got[:, 7]) + 1e-300
    ratio = np.abs(mass) / scale
    assert ratio.max() < 1e-4, ratio.max()
    assert (ratio <= 1e-9).mean() > 0.95, (ratio <= 1e-9).mean()
    sums = model.moment_sums(du)
    assert np.allclose(sums, got.sum(axis=0), rtol=1e-9, atol=1e-9 * np.abs(got).sum(axis=0).max())
    opar = oracle_params(par)
    for i in range(0, n, n // 12):
        ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
        ok, worst = tendency_close(got[i], ref, sc, RTOL)
        assert ok, (i, worst)


def test_wide_parameter_sweep(cb):
    """Gamma + Exponential, threshold 0.5: shape k over [1e-3, 10] (incl. both clamps), scale θ over eight decades, so that
    x_th/θ spans 5e-5 ... 5e3: Taylor zone, series zone, continued-fraction zone and their borders."""
    from cloudy_b200 import workloads as W
    par, _ = W.c2_gamma_exp(n_parcels=8)
    rng = np.random.default_rng(2718)
    n = 600
    k = np.exp(rng.uniform(np.log(1e-3), np.log(10.0), n))
    k[:20] = 10.0
    th = np.exp(rng.uniform(np.log(1e-4), np.log(1e4), n))
    th[20:60] = 0.5 / rng.uniform(17.0, 27.0, 40)      # x_th/θ around the series limits 18..26
    nn = np.exp(rng.uniform(np.log(1e-3), np.log(1e3), n))
    m1 = np.stack([nn, nn * k * th, nn * k * (k + 1) * th ** 2], axis=1)
    n2 = np.exp(rng.uniform(np.log(1e-6), 0.0, n)); th2 = np.exp(rng.uniform(0.0, np.log(30.0), n))
    state = np.concatenate([m1, np.stack([n2, n2 * th2], axis=1)], axis=1) * np.array([1e6, 1e-3, 1e-12, 1e6, 1e-3])
    _check_box(cb, par, state, n, lanes=(0, 1, 8))
    # dense sampling of the regime borders: x_th/θ from just below each series limit (18..26, by shape) to where even the
    # lowest near node leaves the series regime (limit / 0.9), for shapes across all table columns
    kk = np.repeat(np.array([0.05, 0.7, 1.0, 2.3, 3.9, 5.5, 7.2, 8.6, 10.0]), 40)
    lim = np.array([18, 18, 19, 19, 20, 20, 21, 21, 22, 22, 23, 23, 24, 25, 25, 26, 26, 26], dtype=float)[np.floor(kk + 3).astype(int)]
    X = lim - 1.0 + rng.uniform(0.0, 1.0, kk.size) * (lim / 0.9 + 2.0 - lim)
    thb = 0.5 / X
    nb = np.exp(rng.uniform(np.log(1e-1), np.log(1e2), kk.size))
    m1 = np.stack([nb, nb * kk * thb, nb * kk * (kk + 1) * thb ** 2], axis=1)
    n2 = np.full(kk.size, 0.1); th2 = np.full(kk.size, 5.0)
    state = np.concatenate([m1, np.stack([n2, n2 * th2], axis=1)], axis=1) * np.array([1e6, 1e-3, 1e-12, 1e6, 1e-3])
    _check_box(cb, par, state, kk.size, lanes=(0, 1, 8))


def test_dynamic_tile_schedule_sizes_and_repeats(cb):
    """The thread-per-parcel kernel draws its tiles from a device counter that is never reset (tpp_kernel.cuh): every ensemble
    size around the tile and block boundaries, launched back to back on one context, must reproduce the same per-parcel
    results bit for bit, and agree with the lane-cooperative kernel (an independent implementation with a static schedule)."""
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_exp(n_parcels=5000)
    model = cb.CoalescenceModel(par)
    model.ctx.set_lanes(8)
    ref = model.coal_tendency_host(state)
    model.ctx.set_lanes(1)
    full = model.coal_tendency_host(state)
    scale = np.abs(ref).max(axis=0)
    assert np.all(np.abs(full - ref) <= 1e-10 * scale)
    for rep in range(3):
        for n in (1, 2, 31, 32, 33, 127, 128, 129, 255, 257, 1023, 4096, 4097, 5000):
            got = model.coal_tendency_host(state[:n])
            assert np.array_equal(got, full[:n]), (rep, n)
    model.ctx.set_lanes(0)


def test_z_sum_tables_against_node_by_node_sums(cb):
    """Small tensors take the lower-order sums Z[p1][p] from polynomial tables in the Gamma shape k (cloudy_config_set); the
    lane-cooperative kernel sums the same terms node by node.  Shapes over the whole clamp range, including both ends."""
    from cloudy_b200 import workloads as W
    from cloudy_b200.workloads import _moments_from_params, _norm_factors, NORMS
    from cloudy_b200 import _lib as L
    par, _ = W.c2_gamma_exp(n_parcels=8)
    rng = np.random.default_rng(77)
    n = 4096
    k1 = np.concatenate([np.linspace(1e-3, 10.0, n - 6), [1e-6, 1e-4, 9.999, 10.0, 10.0, 0.5]])
    n1 = np.exp(rng.uniform(math.log(1e1), math.log(1e3), n)); th1 = np.exp(rng.uniform(math.log(0.01), math.log(1.0), n))
    n2 = np.exp(rng.uniform(math.log(1e-6), 0.0, n)); th2 = np.exp(rng.uniform(0.0, math.log(30.0), n))
    m = np.concatenate([_moments_from_params(L.GAMMA, n1, th1, k1, 3), _moments_from_params(L.EXPONENTIAL, n2, th2, None, 2)], axis=1)
    state = m * _norm_factors((3, 2), NORMS)
    model = cb.CoalescenceModel(par)
    model.ctx.set_lanes(8)
    ref = model.coal_tendency_host(state)
    model.ctx.set_lanes(1)
    got = model.coal_tendency_host(state)
    model.ctx.set_lanes(0)
    opar = oracle_params(par)
    worst = 0.0
    for i in list(range(0, n, 97)) + list(range(n - 6, n)):
        r, sc = O.rhs_coal(state[i], opar, return_scale=True)
        ok, w = tendency_close(got[i], r, sc, RTOL)
        assert ok, (i, k1[i], w)
        ok2, w2 = tendency_close(got[i], ref[i], sc, 1e-11)  # the two kernels share the rule: two orders tighter
        assert ok2, (i, k1[i], w2)
        worst = max(worst, w2)


def test_given_gamma_shape_outside_the_tables_is_refused(cb):
    dist = (cb.GammaPrimitiveParticleDistribution(100.0, 0.1, 12.0), cb.ExponentialPrimitiveParticleDistribution(1.0, 1.0))
    kernel = cb.CoalescenceTensor(np.array([[0.0, 5e-3], [5e-3, 0.0]]))
    coal_data = cb.CoalescenceData(kernel, (3, 2), (0.5, math.inf))
    with pytest.raises(Exception, match="shape parameter"):
        cb.get_coal_ints(cb.AnalyticalCoalStyle(), dist, coal_data)
