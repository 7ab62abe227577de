"""Error bound of the Z-sum tables (cloudy_config_set): the node-only functions
    G_{p1,p}(k) = sum_j w_j x_j^p1 (x_th - x_j)^p exp(k (ls_j - L)),   ls_j = ln x_j + ln(x_th - x_j),  L = max_j ls_j
of the reference's log-spaced grid (ParticleDistributions.jl:579-585) are mixtures of decaying exponentials in k; the library
tabulates them as degree-7 polynomials on 1024 intervals of [0, 11] (interpolation at Chebyshev nodes, long double).  This
restates that construction in numpy and checks it against direct summation — no GPU, no library call."""
import numpy as np

LD = np.longdouble
N_INT, K_MAX = 1024, 11.0


def _grid(T, nb):
    x_lb = min(1e-5, 1e-5 * T)
    ell = np.log(x_lb) + np.arange(nb) * (np.log(T) - np.log(x_lb)) / nb
    x = np.exp(ell)
    tmx = T - x
    ls = ell + np.log(tmx)
    return x, tmx, (ls - ls.max()).astype(LD)


def _fit(c, d, iv):
    """monomial coefficients (double) of the degree-7 interpolant of sum_j c_j exp(k d_j) on interval iv, in t in [-1, 1]"""
    q = np.arange(8)
    tq = np.cos(np.pi * (q + 0.5) / 8).astype(LD)
    h = LD(K_MAX) / N_INT
    f = np.array([np.sum(c * np.exp(((iv + 0.5 * (t + 1)) * h) * d)) for t in tq])
    a = np.array([np.sum(f * np.cos(m * np.pi * (q + 0.5) / 8)) * (0.125 if m == 0 else 0.25) for m in range(8)])
    Tm = np.zeros((8, 8), dtype=LD)
    Tm[0, 0] = 1
    Tm[1, 1] = 1
    for m in range(1, 7):
        Tm[m + 1, 1:] = 2 * Tm[m, :-1]
        Tm[m + 1] -= Tm[m - 1]
    return (a[:, None] * Tm).sum(0).astype(np.float64)


def test_degree7_tables_reproduce_the_node_sums():
    worst = 0.0
    for T, nb in ((0.5, 75), (100.0, 105), (1e-3, 75)):
        x, tmx, d = _grid(T, nb)
        for p1, p in ((0, 0), (0, 2), (1, 1), (2, 2)):
            c = (x ** p1 * tmx ** p).astype(LD)
            for iv in (0, 1, 7, 300, 1023):
                mono = _fit(c, d, iv)
                for t in np.linspace(-1.0, 1.0, 17):
                    k = (iv + 0.5 * (LD(t) + 1)) * LD(K_MAX) / N_INT
                    exact = np.sum(c * np.exp(k * d))
                    g = mono[7]
                    for cc in mono[6::-1]:
                        g = g * t + cc  # the kernel's Horner (without the fused rounding)
                    worst = max(worst, abs(float((g - exact) / exact)))
    assert worst < 3e-15, worst
