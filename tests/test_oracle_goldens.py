"""The oracle against every golden / known-answer value the reference's own tests hold for the
hot path.  Citations: /root/reference/test/unit_tests/*.jl."""
import math

import numpy as np
import pytest
from scipy import special as sp

from oracle import cloudy_oracle as O

RTOL = 1e-3  # the reference's own tolerance (test_ParticleDistributions_correctness.jl:15)


def approx(a, b, rtol=RTOL):
    return abs(a - b) <= rtol * max(abs(a), abs(b))


# ---- test_helper_functions.jl:5-24 -------------------------------------------------------------
def test_layout_helpers():
    npm = (2, 2, 3)
    assert O.get_dist_moment_ind(npm, 1, 2) == 2
    assert O.get_dist_moment_ind(npm, 2, 1) == 3
    assert O.get_dist_moment_ind(npm, 3, 2) == 6
    for args in ((4, 2), (2, 0), (3, 4)):
        with pytest.raises(Exception):
            O.get_dist_moment_ind(npm, *args)
    assert O.get_dist_moments_ind_range(npm, 1) == range(1, 3)
    assert O.get_dist_moments_ind_range(npm, 3) == range(5, 8)
    with pytest.raises(Exception):
        O.get_dist_moments_ind_range(npm, 4)
    nf = O.get_moments_normalizing_factors(npm, (10.0, 0.1))
    for a, b in zip(nf, (10.0, 1.0, 10.0, 1.0, 10.0, 1.0, 0.1)):
        assert abs(a - b) < 1e-12


# ---- test_KernelTensors_correctness.jl:49-52 ---------------------------------------------------
def test_tensor_normalisation():
    c = np.array([[1.0, 2.0], [2.0, 3.0]])
    assert np.allclose(O.get_normalized_kernel_tensor(c, (10.0, 0.2)), [[10.0, 4.0], [4.0, 1.2]], atol=1e-12, rtol=0)
    with pytest.raises(Exception):
        O.check_symmetry(np.array([[1.0, -0.2], [0.2, 2.0]]))


# ---- test_ParticleDistributions_correctness.jl --------------------------------------------------
def test_constructors_throw():
    for ctor, bad in ((O.Monodisperse, [(-1.0, 2.0), (1.0, -2.0)]), (O.Exponential, [(-1.0, 2.0), (1.0, -2.0)]),
                      (O.Gamma, [(-1.0, 2.0, 3.0), (1.0, -2.0, 3.0), (1.0, 2.0, -3.0)]),
                      (O.Lognormal, [(-1.0, 2.0, 3.0), (1.0, 2.0, -3.0)])):
        for args in bad:
            with pytest.raises(Exception):
                ctor(*args)


def test_moments_exact():
    d = O.Monodisperse(1.0, 2.0)  # :33-37
    assert O.moment(d, 1.0) == 2.0 and O.moment(d, 0.0) == 1.0 and O.moment(d, 10.0) == 2.0 ** 10.0
    d = O.Exponential(1.0, 2.0)  # :70-75
    assert O.moment(d, 1.0) == 2.0 and O.moment(d, 0.0) == 1.0
    assert O.get_moments(d) == [1.0, 2.0]
    assert O.moment(d, 10.0) == 2.0 ** 10.0 * sp.gamma(11.0)
    assert O.density(d, 3.1) == 0.5 * math.exp(-3.1 / 2.0) and O.density(d, 0.0) == 0.5
    d = O.Gamma(1.0, 1.0, 2.0)  # :110-119
    assert [O.moment(d, q) for q in (0.0, 1.0, 2.0)] == [1.0, 2.0, 6.0]
    assert O.get_moments(d) == [1.0, 2.0, 6.0]
    assert approx(O.moment(d, 2 / 3), sp.gamma(2 + 2 / 3) / sp.gamma(2), 1e-14)
    assert O.density(d, 3.0) == 3 / sp.gamma(2) * math.exp(-3)
    d = O.Lognormal(1.0, 1.0, 2.0)  # :156-166
    assert O.moment(d, 0.0) == 1.0 and O.moment(d, 1.0) == math.exp(3.0) and O.moment(d, 2.0) == math.exp(10.0)
    assert approx(O.moment(d, 0.5), math.exp(1.0), 1e-14)
    assert math.isnan(O.density(d, 0.0))
    with pytest.raises(Exception):
        O.density(d, -0.1)


def test_update_dist_from_moments():
    d = O.update_dist_from_moments(O.Monodisperse(1.0, 2.0), (1.1, 2.0))  # :46-54
    assert approx(O.moment(d, 0.0), 1.1) and approx(O.moment(d, 1.0), 2.0)
    d = O.update_dist_from_moments(d, (1.1, 0.0))
    assert O.moment(d, 0.0) == 0.0 and O.moment(d, 1.0) == 0.0
    d = O.update_dist_from_moments(O.Exponential(1.0, 2.0), (10.0, 50.0))  # :89-91
    assert (d.n, d.theta) == (10.0, 5.0)
    with pytest.raises(Exception):
        O.update_dist_from_moments(d, (10.0, 50.0, 300.0))
    g = O.Gamma(1.0, 1.0, 2.0)
    d = O.update_dist_from_moments(g, (1.1, 2.0, 4.1), k_range=(O.EPS, 5.0))  # :128-136
    assert approx(O.moment(d, 0.0), 1.1) and approx(O.moment(d, 1.0), 2.0) and approx(O.moment(d, 2.0), 4.364)
    d = O.update_dist_from_moments(g, (1.1, 2.423, 8.112))
    assert approx(O.moment(d, 2.0), 8.112)
    d = O.update_dist_from_moments(g, (10.0, 50.0, 300.0))  # :141-143
    assert (d.n, d.k, d.theta) == (10.0, 5.0, 1.0)
    with pytest.raises(Exception):
        O.update_dist_from_moments(d, (10.0, 50.0))
    ln = O.Lognormal(1.0, 1.0, 2.0)
    d = O.update_dist_from_moments(ln, (1.1, 2.0, 4.1), mu_range=(-1e5, 1e5), sigma_range=(O.EPS, 5.0))  # :174-181
    assert approx(O.moment(d, 0.0), 1.1) and approx(O.moment(d, 1.0), 2.0) and approx(O.moment(d, 2.0), 4.1)
    d = O.update_dist_from_moments(ln, (10.0, 50.0, 300.0))  # :187-191
    assert approx(d.n, 10.0) and approx(d.mu, 1.518) and approx(d.sigma, 0.427)
    d = O.update_dist_from_moments(ln, (1.1, 0.0, 8.112))  # :196-199
    assert O.moment(d, 0.0) == 0.0 and O.moment(d, 2.0) == 0.0


def test_moment_source_helper_goldens():
    d = O.Monodisperse(1.0, 0.5)  # :201-206
    assert O.moment_source_helper(d, 0.0, 0.0, 0.5) == 0.0
    assert O.moment_source_helper(d, 0.0, 0.0, 1.2) == 1.0
    assert O.moment_source_helper(d, 1.0, 0.0, 0.5) == 0.0
    assert O.moment_source_helper(d, 0.0, 1.0, 1.2) == 0.5
    d = O.Exponential(1.0, 0.5)  # :207-210
    assert approx(O.moment_source_helper(d, 0.0, 0.0, 0.5, 20), 2.642e-1)
    assert approx(O.moment_source_helper(d, 1.0, 0.0, 0.5, 20), 4.015e-2)
    assert approx(O.moment_source_helper(d, 1.0, 1.0, 0.5, 20), 4.748e-3)
    d = O.Gamma(1.0, 0.5, 2.0)  # :211-214
    assert approx(O.moment_source_helper(d, 0.0, 0.0, 0.5, 20), 1.899e-2)
    assert approx(O.moment_source_helper(d, 1.0, 0.0, 0.5, 20), 3.662e-3)
    assert approx(O.moment_source_helper(d, 1.0, 1.0, 0.5, 20), 5.940e-4)


def test_moment_source_helper_lognormal_goldens():
    d = O.Lognormal(1.0, 0.5, 2.0)  # :215-218
    for (p1, p2, gold) in ((0.0, 0.0, 2.831e-1), (1.0, 0.0, 1.725e-1), (1.0, 1.0, 8.115e-2)):
        assert approx(O.moment_source_helper(d, p1, p2, 2.5), gold)
        assert approx(O.moment_source_helper_lognormal_closed(d, p1, p2, 2.5, order=400), gold)


def test_simpson_kat():
    npt = 90  # :250-255
    x = np.linspace(1.0, 10.0, npt + 1)
    dx = x[1] - x[0]
    assert abs(O.integrate_SimpsonEvenFast(npt, dx, lambda j: x[j - 1] ** 2) - 333.0) < 1e-6
    with pytest.raises(Exception):
        O.integrate_SimpsonEvenFast(2, dx, lambda j: 0.0)


def test_compute_thresholds():
    pd = (O.Exponential(10.0, 1.0), O.Gamma(5.0, 10.0, 2.0))  # :257-268
    assert O.compute_threshold(pd[0], 0.75) > 1.0
    assert O.compute_threshold(pd[1], 0.75) > 20.0
    assert abs(O.compute_threshold(pd[0], 0.0)) < 1e-6 and abs(O.compute_threshold(pd[1], 0.0)) < 1e-6
    assert approx(O.compute_thresholds(pd)[0], 3.507)
    assert O.compute_thresholds(pd)[1] > 1e6
    assert approx(O.compute_thresholds(pd, (0.5, 1.0))[0], 0.6931)


# ---- test_Sources_correctness.jl ----------------------------------------------------------------
def test_smoluchowski_1916():
    """:41-85 — constant kernel, M0(t) = 1/(1/a + b t/2), M1 = 2."""
    cd = O.make_coalescence_data(np.array([[1.0]]), (2,), (O.INF,))
    mom = (1.0, 2.0)
    dist = O.Exponential(1.0, 1.0)
    dt = 1e-4
    for _ in range(5):
        ld = O.update_dist_from_moments(dist, mom)
        dm = O.get_coal_ints((ld,), cd)
        mom = tuple(dt * dm[i] + mom[i] for i in range(2))
        dist = ld
    for i in range(6):
        t = dt * i
        ana = 1 / (1 / 1 + 1 / 2 * t)
        assert approx(mom[0], ana) and approx(mom[1], 2.0)
    assert abs(mom[0] - 0.9997500499912514) < 1e-15  # SURVEY Appendix B restatement KAT


def _gamma_exp_setup():
    dist = (O.Gamma(100.0, 0.1, 1.0), O.Exponential(1.0, 1.0))
    c = np.array([[0.0, 5e-3], [5e-3, 0.0]])  # exact LinearKernelFunction(5e-3) tensor
    cd = O.make_coalescence_data(c, (3, 2), (0.5, O.INF))
    return dist, c, cd


def test_gamma_exp_independent_restatement():
    """:89-169 — get_coal_ints equals the test's own loop restatement to 10 eps."""
    dist, c, cd = _gamma_exp_setup()
    npm = (3, 2)
    order = 1
    ci = O.get_coal_ints(dist, cd)
    n_mom = max(npm) + order
    mom = np.array([[O.moment(dist[i], float(j)) for j in range(n_mom)] for i in range(2)])
    iw = np.zeros((n_mom, n_mom))
    mm = np.zeros((n_mom, n_mom))
    for i in range(n_mom):
        for j in range(i, n_mom):
            mm[i, j] = mom[0, i] * mom[0, j]
            tmp = 0.0 if mm[i, j] < O.EPS else O.moment_source_helper(dist[0], float(i), float(j), 0.5)
            iw[i, j] = min(mm[i, j], tmp)
            mm[j, i] = mm[i, j]
            iw[j, i] = iw[i, j]
    out = np.zeros(5)
    for i in range(2):
        j = 1 if i == 0 else 0
        for k in range(npm[i]):
            temp = 0.0
            for a in range(order + 1):
                for b in range(order + 1):
                    coef = c[a, b]
                    temp -= coef * mom[i, a + k] * mom[i, b]
                    temp -= coef * mom[i, a + k] * mom[j, b]
                    for cc in range(k + 1):
                        cb = coef * math.comb(k, cc)
                        if i == 0:
                            temp += 0.5 * cb * iw[a + cc, b + k - cc]
                        else:
                            temp += 0.5 * cb * (mm[a + cc, b + k - cc] - iw[a + cc, b + k - cc])
                            temp += 0.5 * cb * mom[i, a + cc] * mom[i, b + k - cc]
                            temp += cb * mom[j, a + cc] * mom[i, b + k - cc]
            out[O.get_dist_moment_ind(npm, i + 1, k + 1) - 1] = temp
    # the reference asserts rtol 10 eps on entries 1,3,4,5 (:158-169); summation order differs here, allow 1e-13
    assert np.allclose(ci, out, rtol=1e-13, atol=1e-13 * np.abs(out).max())


def test_gamma_exp_invariants_and_kat():
    dist, c, cd = _gamma_exp_setup()
    ci = O.get_coal_ints(dist, cd)
    # SURVEY Appendix B KAT-A (restatement KAT, exact linear tensor)
    kat = (-6.1734944299460315, -0.45756875772176486, -0.0773368945825677, 0.6184944299460307, 0.457568757721765)
    assert np.allclose(ci, kat, rtol=1e-12, atol=0)
    assert abs(ci[0] + ci[3] + 5.555) < 1e-12  # Σ number tendency = -c12 M0_tot M1_tot
    assert abs(ci[1] + ci[4]) < 1e-14  # mass conserved


def test_sedimentation_golden():
    pd = (O.Exponential(1.0, 1.0),)  # :267-272
    flux = O.get_sedimentation_flux(pd, ((1.0, 0.0), (-1.0, 1.0 / 6)))
    assert np.allclose(flux, (-1.0 + sp.gamma(1.0 + 1.0 / 6), -1.0 + sp.gamma(2.0 + 1.0 / 6)), rtol=1e-15)


def test_condensation_golden():
    pd = (O.Exponential(1.0, 1.0),)  # :276-285
    got = O.get_cond_evap(pd, 0.01, 1e-6)
    want = (0.0, 3 * 1e-6 * 0.01 * O.moment(pd[0], 1 - 2 / 3) * (4 * math.pi / 3) ** (2 / 3) / 1000.0 ** (1 / 3))
    assert np.allclose(got, want, rtol=1e-15)


def test_box_kat_d():
    """SURVEY Appendix B KAT-D: box_gamma_mixture.jl:14-36 initial tendency (restatement KAT)."""
    c = np.array([[0.0, 5.0], [5.0, 0.0]])
    cd = O.make_coalescence_data(c, (3, 3), (5e-10, O.INF), (1e6, 1e-9))
    par = O.ModelParams((O.GAMMA, O.GAMMA), cd, (3, 3), (1e6, 1e-9))
    dm = O.rhs_coal([1e8, 1e-2, 2e-12, 1, 1e-8, 2e-16], par)
    kat = (-5623499.479946031, -3.975692677217648e-4, -6.433699758256767e-14, 623494.4299459805,
           3.975692677217648e-4, 2.6435719760256764e-13)
    assert np.allclose(dm, kat, rtol=1e-11, atol=0)


def test_mpmath_crosscheck():
    """The double-precision rule against a 50-digit evaluation of the SAME rule (same nodes)."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rng = np.random.default_rng(7)
    worst = 0.0
    for _ in range(6):
        k = float(rng.uniform(0.3, 8.0))
        theta = float(10 ** rng.uniform(-1.5, 0.5))
        T = float(10 ** rng.uniform(-1, 1))
        p1, p2 = sorted(int(v) for v in rng.integers(0, 4, size=2))
        d = O.Gamma(3.0, theta, k)
        got = O.moment_source_helper(d, float(p1), float(p2), T)
        n_bins, x_min, dx = O.log_grid(T)
        ys = {}

        def y(j):
            if j > n_bins:
                return mp.mpf(0)
            if j not in ys:
                x = mp.mpf(math.exp(x_min + (j - 1) * dx))
                ys[j] = x * x ** (p1 + mp.mpf(k) - 1) * mp.exp(-x / mp.mpf(theta)) * \
                    mp.gammainc(p2 + mp.mpf(k), 0, (mp.mpf(T) - x) / mp.mpf(theta))
            return ys[j]

        e = n_bins + 1
        s = sum(y(j) for j in range(5, n_bins - 2))
        s += (17 * (y(1) + y(e)) + 59 * (y(2) + y(e - 1)) + 43 * (y(3) + y(e - 2)) + 49 * (y(4) + y(e - 3))) / 48
        want = mp.mpf(3.0) ** 2 * mp.mpf(theta) ** (p2 - mp.mpf(k)) / mp.gamma(mp.mpf(k)) ** 2 * mp.mpf(dx) * s
        worst = max(worst, abs(float((mp.mpf(got) - want) / want)))
    assert worst < 2e-13, worst
