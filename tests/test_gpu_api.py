"""The reference's own unit tests for the path, run through the host mirror (which calls the C ABI / CUDA).
Reads like /root/reference/test/unit_tests/test_ParticleDistributions_correctness.jl and test_Sources_correctness.jl."""
import math

import numpy as np
import pytest
from scipy.special import gamma

pytestmark = pytest.mark.gpu
rtol = 1e-3


def approx(a, b, rtol=rtol):
    return abs(a - b) <= rtol * max(abs(a), abs(b))


@pytest.fixture(scope="module")
def cb():
    import cloudy_b200
    return cloudy_b200


def test_distributions(cb):
    Mono, Exp, Gam, LogN = (cb.MonodispersePrimitiveParticleDistribution, cb.ExponentialPrimitiveParticleDistribution,
                            cb.GammaPrimitiveParticleDistribution, cb.LognormalPrimitiveParticleDistribution)
    for ctor, bad in ((Mono, [(-1.0, 2.0), (1.0, -2.0)]), (Exp, [(-1.0, 2.0), (1.0, -2.0)]),
                      (Gam, [(-1.0, 2.0, 3.0), (1.0, -2.0, 3.0), (1.0, 2.0, -3.0)]), (LogN, [(-1.0, 2.0, 3.0), (1.0, 2.0, -3.0)])):
        for a in bad:
            with pytest.raises(Exception):
                ctor(*a)
    d = Mono(1.0, 2.0)
    assert cb.nparams(d) == 2
    assert cb.moment(d, 1.0) == 2.0 and cb.moment(d, 0.0) == 1.0 and approx(cb.moment(d, 10.0), 2.0 ** 10, 1e-14)
    d = Exp(1.0, 2.0)
    assert cb.moment(d, 1.0) == 2.0 and cb.moment(d, 0.0) == 1.0
    assert approx(cb.moment(d, 10.0), 2.0 ** 10.0 * gamma(11.0), 1e-13)
    assert cb.get_moments(d) == [1.0, 2.0]
    assert cb.density(d, 3.1) == 0.5 * math.exp(-3.1 / 2.0)
    d = Gam(1.0, 1.0, 2.0)
    assert cb.nparams(d) == 3
    assert np.allclose(cb.get_moments(d), [1.0, 2.0, 6.0], rtol=1e-14)
    assert approx(cb.moment(d, 2 / 3), gamma(2 + 2 / 3) / gamma(2), 1e-13)
    d = LogN(1.0, 1.0, 2.0)
    assert approx(cb.moment(d, 1.0), math.exp(3.0), 1e-14) and approx(cb.moment(d, 2.0), math.exp(10.0), 1e-14)
    assert approx(cb.moment(d, 0.5), math.exp(1.0), 1e-14)


def test_update_dist_from_moments(cb):
    Mono, Exp, Gam, LogN = (cb.MonodispersePrimitiveParticleDistribution, cb.ExponentialPrimitiveParticleDistribution,
                            cb.GammaPrimitiveParticleDistribution, cb.LognormalPrimitiveParticleDistribution)
    d = cb.update_dist_from_moments(Mono(1.0, 2.0), (1.1, 2.0))
    assert approx(cb.moment(d, 0.0), 1.1) and approx(cb.moment(d, 1.0), 2.0)
    d = cb.update_dist_from_moments(d, (1.1, 0.0))
    assert cb.moment(d, 0.0) == 0.0 and cb.moment(d, 1.0) == 0.0
    d = cb.update_dist_from_moments(Exp(1.0, 2.0), (10.0, 50.0))
    assert (d.n, d.θ) == (10.0, 5.0)
    with pytest.raises(Exception):
        cb.update_dist_from_moments(d, (10.0, 50.0, 300.0))
    g = Gam(1.0, 1.0, 2.0)
    d = cb.update_dist_from_moments(g, (1.1, 2.0, 4.1), param_range={"k": (2.2e-16, 5.0)})
    assert approx(cb.moment(d, 0.0), 1.1) and approx(cb.moment(d, 1.0), 2.0) and approx(cb.moment(d, 2.0), 4.364)
    d = cb.update_dist_from_moments(g, (1.1, 2.423, 8.112))
    assert approx(cb.moment(d, 2.0), 8.112)
    d = cb.update_dist_from_moments(g, (10.0, 50.0, 300.0))
    assert (d.n, d.k, d.θ) == (10.0, 5.0, 1.0)
    with pytest.raises(Exception):
        cb.update_dist_from_moments(d, (10.0, 50.0))
    ln = LogN(1.0, 1.0, 2.0)
    d = cb.update_dist_from_moments(ln, (1.1, 2.0, 4.1), param_range={"μ": (-1e5, 1e5), "σ": (2.2e-16, 5.0)})
    assert approx(cb.moment(d, 0.0), 1.1) and approx(cb.moment(d, 1.0), 2.0) and approx(cb.moment(d, 2.0), 4.1)
    d = cb.update_dist_from_moments(ln, (10.0, 50.0, 300.0))
    assert approx(d.n, 10.0) and approx(d.μ, 1.518) and approx(d.σ, 0.427)
    d = cb.update_dist_from_moments(ln, (1.1, 0.0, 8.112))
    assert cb.moment(d, 0.0) == 0.0
    with pytest.raises(Exception):  # DomainError in the reference (sqrt(log(x)) with x < 1)
        cb.update_dist_from_moments(ln, (10.0, 50.0, 200.0))


def test_moment_source_helper_goldens(cb):
    from oracle import cloudy_oracle as O
    Mono, Exp, Gam = (cb.MonodispersePrimitiveParticleDistribution, cb.ExponentialPrimitiveParticleDistribution,
                      cb.GammaPrimitiveParticleDistribution)
    d = Mono(1.0, 0.5)
    assert cb.moment_source_helper(d, 0.0, 0.0, 0.5) == 0.0
    assert cb.moment_source_helper(d, 0.0, 0.0, 1.2) == 1.0
    assert cb.moment_source_helper(d, 1.0, 0.0, 0.5) == 0.0
    assert cb.moment_source_helper(d, 0.0, 1.0, 1.2) == 0.5
    d = Exp(1.0, 0.5)
    assert approx(cb.moment_source_helper(d, 0.0, 0.0, 0.5, 20), 2.642e-1)
    assert approx(cb.moment_source_helper(d, 1.0, 0.0, 0.5, 20), 4.015e-2)
    assert approx(cb.moment_source_helper(d, 1.0, 1.0, 0.5, 20), 4.748e-3)
    d = Gam(1.0, 0.5, 2.0)
    assert approx(cb.moment_source_helper(d, 0.0, 0.0, 0.5, 20), 1.899e-2)
    assert approx(cb.moment_source_helper(d, 1.0, 0.0, 0.5, 20), 3.662e-3)
    assert approx(cb.moment_source_helper(d, 1.0, 1.0, 0.5, 20), 5.940e-4)
    # and to 1e-12 against the oracle, including real-valued orders and the performance-test arguments
    for (dd, od, p1, p2, xt, nb) in ((Gam(5.0, 10.0, 2.0), O.Gamma(5.0, 10.0, 2.0), 1.0, 0.0, 1.2, 15),
                                     (Exp(10.0, 1.0), O.Exponential(10.0, 1.0), 1.0, 0.0, 1.2, 15),
                                     (Gam(100.0, 0.1, 1.0), O.Gamma(100.0, 0.1, 1.0), 2.0, 3.0, 0.5, 15),
                                     (Gam(3.0, 0.2, 0.7), O.Gamma(3.0, 0.2, 0.7), 0.5, 1.5, 40.0, 15)):
        assert approx(cb.moment_source_helper(dd, p1, p2, xt, nb), O.moment_source_helper(od, p1, p2, xt, nb), 1e-12)


def test_simpson_kat(cb):
    npt = 90
    x = np.linspace(1.0, 10.0, npt + 1)
    dx = x[1] - x[0]
    assert abs(cb.integrate_SimpsonEvenFast(npt, dx, lambda j: x[j - 1] ** 2) - 333.0) < 1e-6
    with pytest.raises(Exception):
        cb.integrate_SimpsonEvenFast(2, dx, lambda j: 0.0)


def test_smoluchowski_1916(cb):
    ker = cb.CoalescenceTensor(np.array([[1.0]]))
    mom = (1.0, 2.0)
    dist = (cb.ExponentialPrimitiveParticleDistribution(1.0, 1.0),)
    coal_data = cb.CoalescenceData(ker, (cb.nparams(dist[0]),), (math.inf,))
    dt = 1e-4
    for _ in range(5):
        ldist = (cb.update_dist_from_moments(dist[0], mom),)
        dmom = cb.get_coal_ints(cb.AnalyticalCoalStyle(), ldist, coal_data)
        mom = tuple(dt * dmom[i] + mom[i] for i in range(2))
        dist = ldist
    for i in range(6):
        assert approx(mom[0], 1 / (1 + 0.5 * dt * i)) and approx(mom[1], 2.0)
    assert abs(mom[0] - 0.9997500499912514) < 1e-14


def test_gamma_exp_get_coal_ints(cb):
    """test_Sources_correctness.jl:89-169 with the exact linear tensor; SURVEY Appendix B KAT-A."""
    dist = (cb.GammaPrimitiveParticleDistribution(100.0, 0.1, 1.0), cb.ExponentialPrimitiveParticleDistribution(1.0, 1.0))
    kernel = cb.CoalescenceTensor(np.array([[0.0, 5e-3], [5e-3, 0.0]]))
    coal_data = cb.CoalescenceData(kernel, (3, 2), (0.5, math.inf))
    ci = cb.get_coal_ints(cb.AnalyticalCoalStyle(), dist, coal_data)
    kat = (-6.1734944299460315, -0.45756875772176486, -0.0773368945825677, 0.6184944299460307, 0.457568757721765)
    assert np.allclose(ci, kat, rtol=1e-11, atol=0)
    assert abs(ci[0] + ci[3] + 5.555) < 1e-11 and abs(ci[1] + ci[4]) < 1e-13
    with pytest.raises(Exception):
        cb.get_coal_ints(cb.NumericalCoalStyle(), dist, coal_data)


def test_sedimentation_golden(cb):
    pd = (cb.ExponentialPrimitiveParticleDistribution(1.0, 1.0),)
    flux = cb.get_sedimentation_flux(pd, ((1.0, 0.0), (-1.0, 1.0 / 6)))
    assert np.allclose(flux, (-1.0 + gamma(1.0 + 1.0 / 6), -1.0 + gamma(2.0 + 1.0 / 6)), rtol=1e-14)


def test_box_model_rhs_callable(cb):
    """make_box_model_rhs: the reference's in-place rhs!(dm, m, par, t) on one moment vector."""
    from cloudy_b200 import workloads as W
    par, state = W.c2_gamma_gamma(n_parcels=4)
    rhs = cb.make_box_model_rhs(cb.AnalyticalCoalStyle())
    dm = np.zeros(6)
    rhs(dm, state[0], par, 0.0)
    kat = (-5623499.479946031, -3.975692677217648e-4, -6.433699758256767e-14, 623494.4299459805,
           3.975692677217648e-4, 2.6435719760256764e-13)
    assert np.allclose(dm, kat, rtol=1e-10, atol=0)
    with pytest.raises(Exception):
        cb.make_box_model_rhs(cb.NumericalCoalStyle())


def test_error_behaviour(cb):
    with pytest.raises(Exception):
        cb.CoalescenceTensor(np.array([[1.0, -0.2], [0.2, 2.0]]))  # array not symmetric.
    with pytest.raises(Exception):
        cb.get_moments_normalizing_factors((2, 2), (0.0, 1.0))  # norms must be positive!
    with pytest.raises(Exception):
        cb.get_dist_moment_ind((2, 2, 3), 4, 2)


def test_condensation_golden(cb):
    """test_Sources_correctness.jl:274-312"""
    Exp, Gam = cb.ExponentialPrimitiveParticleDistribution, cb.GammaPrimitiveParticleDistribution
    ξ, s = 1e-6, 0.01
    geom = (4 * math.pi / 3) ** (2 / 3) / 1000.0 ** (1 / 3)
    pd = (Exp(1.0, 1.0),)
    got = cb.get_cond_evap(pd, s, ξ)
    assert got[0] == 0.0 and approx(got[1], 3 * ξ * s * cb.moment(pd[0], 1 - 2 / 3) * geom, 1e-13)
    pd = (Exp(1.0, 1.0), Gam(1.0, 2.0, 3.0), Gam(0.1, 10.0, 3.0))
    got = cb.get_cond_evap(pd, s, ξ)
    want = (0.0, 3 * ξ * s * cb.moment(pd[0], 1 - 2 / 3) * geom,
            0.0, 3 * ξ * s * cb.moment(pd[1], 1 - 2 / 3) * geom, 3 * 2 * ξ * s * cb.moment(pd[1], 2 - 2 / 3) * geom,
            0.0, 3 * ξ * s * cb.moment(pd[2], 1 - 2 / 3) * geom, 3 * 2 * ξ * s * cb.moment(pd[2], 2 - 2 / 3) * geom)
    assert np.allclose(got, want, rtol=1e-13, atol=0)


def test_get_standard_N_q(cb):
    """test_ParticleDistributions_correctness.jl:234-247"""
    from oracle import cloudy_oracle as O
    Exp, Gam, Mono, LogN = (cb.ExponentialPrimitiveParticleDistribution, cb.GammaPrimitiveParticleDistribution,
                            cb.MonodispersePrimitiveParticleDistribution, cb.LognormalPrimitiveParticleDistribution)
    pd = (Exp(10.0, 1.0), Gam(5.0, 10.0, 2.0))
    q1 = cb.get_standard_N_q(pd, 1.0)
    q2 = cb.get_standard_N_q(pd, 0.5)
    for q in (q1, q2):
        assert approx(q.N_liq + q.N_rai, 15.0) and approx(q.M_liq + q.M_rai, 110.0)
    assert q1.N_liq > q2.N_liq and q1.M_liq > q2.M_liq
    ref = O.get_standard_N_q((O.Exponential(10.0, 1.0), O.Gamma(5.0, 10.0, 2.0)), 1.0)
    assert np.allclose((q1.N_liq, q1.N_rai, q1.M_liq, q1.M_rai), ref, rtol=1e-13)
    # the performance-test tuple (performance_tests.jl:94-100): Monodisperse, Lognormal, Gamma
    pd = (Mono(1.0, 0.5), LogN(1.0, 0.5, 2.0), Gam(5.0, 10.0, 2.0))
    q = cb.get_standard_N_q(pd, 1.2)
    ref = O.get_standard_N_q((O.Monodisperse(1.0, 0.5), O.Lognormal(1.0, 0.5, 2.0), O.Gamma(5.0, 10.0, 2.0)), 1.2)
    assert np.allclose((q.N_liq, q.N_rai, q.M_liq, q.M_rai), ref, rtol=1e-7)  # Lognormal: the reference integrates adaptively at sqrt(eps)


def test_compute_thresholds(cb):
    """test_ParticleDistributions_correctness.jl:257-268"""
    from oracle import cloudy_oracle as O
    pd = (cb.ExponentialPrimitiveParticleDistribution(10.0, 1.0), cb.GammaPrimitiveParticleDistribution(5.0, 10.0, 2.0))
    assert cb.compute_threshold(pd[0], 0.75) > 1.0
    assert cb.compute_threshold(pd[1], 0.75) > 2.0 * 10.0
    assert abs(cb.compute_threshold(pd[0], 0.0)) < 1e-6 and abs(cb.compute_threshold(pd[1], 0.0)) < 1e-6
    assert approx(cb.compute_thresholds(pd)[0], 3.507)
    assert cb.compute_thresholds(pd)[1] > 1e6
    assert approx(cb.compute_thresholds(pd, (0.5, 1.0))[0], 0.6931)
    for k, pct in ((0.3, 0.9), (2.0, 0.97), (7.5, 0.5), (10.0, 0.999), (1.0, 0.01)):
        got = cb.compute_threshold(cb.GammaPrimitiveParticleDistribution(1.0, 3.0, k), pct)
        assert approx(got, O.compute_threshold(O.Gamma(1.0, 3.0, k), pct), 1e-13)
    with pytest.raises(Exception):
        cb.compute_threshold(cb.LognormalPrimitiveParticleDistribution(1.0, 0.0, 1.0), 0.9)


def test_get_coal_ints_moving_threshold(cb):
    """get_coal_ints(cs, pdists, coal_data, MovingThreshold()) — Coalescence.jl:152-185"""
    from oracle import cloudy_oracle as O
    Gam = cb.GammaPrimitiveParticleDistribution
    pd = (Gam(100.0, 0.1, 1.5), Gam(1.0, 8.0, 2.5), Gam(1e-3, 400.0, 1.0))
    ker = cb.CoalescenceTensor(np.array([[0.0, 5e-3], [5e-3, 0.0]]))
    cdm = cb.CoalescenceData(ker, (3, 3, 3), (0.99, 0.95, 1.0), (1.0, 1.0), cb.MovingThreshold())
    got = cb.get_coal_ints(cb.AnalyticalCoalStyle(), pd, cdm, cb.MovingThreshold())
    ocd = O.make_coalescence_data(np.array([[0.0, 5e-3], [5e-3, 0.0]]), (3, 3, 3), (0.99, 0.95, 1.0), (1.0, 1.0), moving=True)
    ref, sc = O.get_coal_ints((O.Gamma(100.0, 0.1, 1.5), O.Gamma(1.0, 8.0, 2.5), O.Gamma(1e-3, 400.0, 1.0)), ocd, True)
    assert np.all(np.abs(np.array(got) - ref) <= 1e-9 * np.maximum(np.abs(ref), sc))
    assert abs(got[1] + got[4] + got[7]) <= 1e-12 * (abs(got[1]) + abs(got[4]) + abs(got[7]))
