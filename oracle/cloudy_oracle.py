"""CPU oracle for the Cloudy.jl coalescence-tendency hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain restatement, in numpy/scipy,
of the reference's Julia algorithm.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it;
the product (``cloudy.jl_b200``) never does and fails loudly without its CUDA
library.

Parity status
-------------
Julia is not installed in the build container, so the reference itself cannot
be executed.  The oracle is pinned against every golden value the reference's
own tests hold for this path (``tests/test_oracle_goldens.py``); those goldens
are 4-significant-digit values (rtol 1e-3) at the ``gamma_inc`` boundary, so at
the 1e-9 level the path is **parity unpinned by the reference's own tests** —
the 1e-9 claim rests on this restatement being a faithful transcription of the
formulas (cross-checked against a 50-digit mpmath evaluation of the same rule
in ``tests/test_oracle_goldens.py::test_mpmath_crosscheck``).

Third-party arithmetic the reference takes from un-vendored packages
(SpecialFunctions.jl compat "2.5": ``gamma``, ``gamma_inc``, ``gamma_inc_inv``;
QuadGK.jl compat "2.11"; OrdinaryDiffEqSSPRK ``SSPRK33``) is replaced by
``scipy.special.{gamma, gammainc, gammaincinv}``, ``scipy.integrate.quad`` and the
published Shu-Osher SSP(3,3) scheme.

All citations are ``path:line`` under ``/root/reference``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Sequence, Tuple

import numpy as np
from scipy import special as sp
from scipy import integrate as spint

EPS = float(np.finfo(np.float64).eps)
INF = float("inf")

EXPONENTIAL, GAMMA, LOGNORMAL, MONODISPERSE = 0, 1, 2, 3
KIND_NAMES = {EXPONENTIAL: "Exponential", GAMMA: "Gamma", LOGNORMAL: "Lognormal", MONODISPERSE: "Monodisperse"}


# --------------------------------------------------------------------------------------
# Julia semantics helpers
# --------------------------------------------------------------------------------------
def jl_min(a: float, b: float) -> float:
    """Julia ``min`` propagates NaN (Python's does not)."""
    if a != a or b != b:
        return float("nan")
    return a if a < b else b


def jl_max(a: float, b: float) -> float:
    if a != a or b != b:
        return float("nan")
    return a if a > b else b


def _div(a: float, b: float) -> float:
    """IEEE division (Julia never raises on x/0)."""
    return float(np.float64(a) / np.float64(b)) if b == 0 else a / b


# --------------------------------------------------------------------------------------
# Distributions — src/ParticleDistributions/ParticleDistributions.jl:66-159
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Dist:
    kind: int
    n: float
    p1: float  # θ (Exp/Gamma/Mono) or μ (Lognormal)
    p2: float = 1.0  # k (Gamma) or σ (Lognormal); unused otherwise

    def __post_init__(self):
        # constructors: ParticleDistributions.jl:72-78, :101-106, :126-131, :153-158
        if self.kind in (EXPONENTIAL, MONODISPERSE):
            if self.n < 0 or self.p1 <= 0:
                raise ValueError("n needs to be nonnegative. θ needs to be positive.")
        elif self.kind == GAMMA:
            if self.n < 0 or self.p1 <= 0 or self.p2 <= 0:
                raise ValueError("n needs to be nonnegative. θ and k need to be positive.")
        elif self.kind == LOGNORMAL:
            if self.n < 0 or self.p2 <= 0:
                raise ValueError("n needs to be nonnegative. σ needs to be positive.")
        else:
            raise ValueError("unknown distribution kind")

    @property
    def theta(self):
        return self.p1

    @property
    def k(self):
        return self.p2

    @property
    def mu(self):
        return self.p1

    @property
    def sigma(self):
        return self.p2


def Exponential(n, theta):
    return Dist(EXPONENTIAL, float(n), float(theta))


def Gamma(n, theta, k):
    return Dist(GAMMA, float(n), float(theta), float(k))


def Lognormal(n, mu, sigma):
    return Dist(LOGNORMAL, float(n), float(mu), float(sigma))


def Monodisperse(n, theta):
    return Dist(MONODISPERSE, float(n), float(theta))


def nparams(d: Dist) -> int:
    """ParticleDistributions.jl:425-427 (number of struct fields)."""
    return 3 if d.kind in (GAMMA, LOGNORMAL) else 2


def moment(d: Dist, q: float) -> float:
    """ParticleDistributions.jl:177-207, :216."""
    if d.kind == EXPONENTIAL:
        return d.n * d.p1 ** q * float(sp.gamma(q + 1.0))
    if d.kind == GAMMA:
        return d.n * d.p1 ** q * float(sp.gamma(q + d.p2)) / float(sp.gamma(d.p2))
    if d.kind == MONODISPERSE:
        return d.n * d.p1 ** q
    return d.n * math.exp(q * d.p1 + q ** 2 * d.p2 ** 2 / 2)


def get_moments(d: Dist):
    """ParticleDistributions.jl:293-315."""
    if d.kind == GAMMA:
        return [d.n, d.n * d.p2 * d.p1, d.n * d.p2 * (d.p2 + 1) * d.p1 ** 2]
    if d.kind == LOGNORMAL:
        return [d.n, d.n * math.exp(d.p1 + d.p2 ** 2 / 2), d.n * math.exp(2.0 * d.p1 + 2.0 * d.p2 ** 2)]
    return [d.n, d.n * d.p1]


def density(d: Dist, x: float) -> float:
    """ParticleDistributions.jl:323-355, :397-402."""
    if x < 0:
        raise ValueError("Density can only be evaluated at nonnegative values.")
    if d.kind == EXPONENTIAL:
        return d.n / d.p1 * math.exp(-x / d.p1)
    if d.kind == GAMMA:
        return d.n * x ** (d.p2 - 1) / d.p1 ** d.p2 / float(sp.gamma(d.p2)) * math.exp(-x / d.p1)
    if d.kind == LOGNORMAL:
        if x == 0:
            return float("nan")
        return d.n * math.exp(-((math.log(x) - d.p1) ** 2 / (2 * d.p2 ** 2))) / (x * d.p2 * math.sqrt(2 * math.pi))
    return d.n / (2 * d.p1 / 10.0) if abs(x - d.p1) < d.p1 / 10.0 else 0.0


def update_dist_from_moments(d: Dist, m: Sequence[float], k_range=(EPS, 10.0),
                             mu_range=(-INF, INF), sigma_range=(EPS, INF)) -> Dist:
    """ParticleDistributions.jl:456-541."""
    if len(m) != nparams(d):
        raise TypeError("wrong number of moments for this distribution")
    if d.kind == GAMMA:
        if m[0] > EPS and m[1] > EPS:
            n = m[0]
            k = jl_max(k_range[0], jl_min(k_range[1], _div(m[1] / m[0], (m[2] / m[1] - m[1] / m[0]))))
            theta = m[1] / m[0] / k
            return Dist(GAMMA, n, theta, k)
        return Dist(GAMMA, 0.0, 1.0, 1.0)
    if d.kind == LOGNORMAL:
        if m[0] > EPS and m[1] > EPS and m[2] > EPS:
            mu = jl_max(mu_range[0], jl_min(mu_range[1], math.log(m[1] ** 2 / m[0] ** (3 / 2) / m[2] ** (1 / 2))))
            arg = math.log(m[0] * m[2] / m[1] ** 2)
            if arg < 0:
                raise ValueError("DomainError: sqrt of negative (ParticleDistributions.jl:498)")
            sigma = jl_max(sigma_range[0], jl_min(sigma_range[1], math.sqrt(arg)))
            n = m[1] / math.exp(mu + 1 / 2 * sigma ** 2)
            return Dist(LOGNORMAL, n, mu, sigma)
        return Dist(LOGNORMAL, 0.0, 1.0, 1.0)
    # Exponential / Monodisperse: :512-541
    if m[0] > EPS and m[1] > EPS:
        return Dist(d.kind, m[0], m[1] / m[0])
    return Dist(d.kind, 0.0, 1.0)


# --------------------------------------------------------------------------------------
# Simpson rule and truncated 2-D integral — ParticleDistributions.jl:557-625, :698-710
# --------------------------------------------------------------------------------------
def integrate_SimpsonEvenFast(n_bins: int, dx: float, y) -> float:
    """ParticleDistributions.jl:698-710; ``y(j)`` is 1-based like the reference."""
    if n_bins < 3:
        raise ValueError("n_bins must be at least 3")
    e = n_bins + 1
    s = 0.0
    for j in range(5, n_bins - 3 + 1):
        s += y(j)
    retval = s + (17 * (y(1) + y(e)) + 59 * (y(2) + y(e - 1)) + 43 * (y(3) + y(e - 2)) + 49 * (y(4) + y(e - 3))) / 48
    return dx * retval


def log_grid(x_threshold: float, n_bins_per_log_unit: int = 15) -> Tuple[int, float, float]:
    """Node grid of ParticleDistributions.jl:579-582 / :604-607 → (n_bins, x_min, dx)."""
    x_lowerbound = min(1e-5, 1e-5 * x_threshold)
    n_bins = int(math.floor(n_bins_per_log_unit * math.log10(x_threshold / x_lowerbound)))
    x_min = math.log(x_lowerbound)
    dx = (math.log(x_threshold) - math.log(x_lowerbound)) / n_bins
    return n_bins, x_min, dx


def _lognormal_density_vec(d: Dist, x):
    return d.n * np.exp(-((np.log(x) - d.p1) ** 2 / (2 * d.p2 ** 2))) / (x * d.p2 * math.sqrt(2 * math.pi))


def moment_source_helper(d: Dist, p1: float, p2: float, x_threshold: float, n_bins_per_log_unit: int = 15) -> float:
    """ParticleDistributions.jl:557-625."""
    if d.kind == MONODISPERSE:  # :557-564
        return d.n ** 2 * d.p1 ** (p1 + p2) if d.p1 < x_threshold / 2 else 0.0
    if d.kind == LOGNORMAL and LOGNORMAL_FIXED_RULE is not None:
        return moment_source_helper_lognormal_closed(d, p1, p2, x_threshold, LOGNORMAL_FIXED_RULE[0], LOGNORMAL_FIXED_RULE[1])
    if d.kind == LOGNORMAL:  # :614-625 — nested adaptive GK (QuadGK default rtol sqrt(eps))
        rt = math.sqrt(EPS)

        def f(y):
            upper = x_threshold - y
            if upper <= 0:
                return 0.0
            inner = spint.quad(lambda xx: xx ** p1 * y ** p2 * float(_lognormal_density_vec(d, xx)) *
                               float(_lognormal_density_vec(d, y)), 0.0, upper, epsabs=0.0, epsrel=rt, limit=200)[0]
            return inner

        return spint.quad(f, 0.0, x_threshold, epsabs=0.0, epsrel=rt, limit=200)[0]
    # Exponential :567-587 is the Gamma formula :589-612 with k = 1
    n, theta = d.n, d.p1
    k = d.p2 if d.kind == GAMMA else 1.0
    gam_p2k = float(sp.gamma(p2 + k))
    n_bins, x_min, dx = log_grid(x_threshold, n_bins_per_log_unit)

    if d.kind == GAMMA:
        def f(x):
            return x ** (p1 + k - 1) * math.exp(-x / theta) * float(sp.gammainc(p2 + k, (x_threshold - x) / theta)) * gam_p2k
    else:
        def f(x):
            return x ** p1 * math.exp(-x / theta) * float(sp.gammainc(p2 + 1, (x_threshold - x) / theta)) * gam_p2k

    def y_func(j):
        if j <= n_bins:
            xj = math.exp(x_min + (j - 1) * dx)
            return xj * f(xj)
        return 0.0

    simpson = integrate_SimpsonEvenFast(n_bins, dx, y_func)
    if d.kind == GAMMA:
        return n ** 2 * theta ** (p2 - k) / float(sp.gamma(k)) ** 2 * simpson
    return n ** 2 * theta ** (p2 - 1) * simpson


# When set to (order, top_order), moment_source_helper evaluates Lognormal modes with the fixed closed-form rule below instead
# of the nested adaptive quadrature: the rule the CUDA path uses (128 points, range up to the mode's top carried order), so
# that the two implementations of the SAME rule can be compared at rounding level (tests only).
LOGNORMAL_FIXED_RULE = None


def moment_source_helper_lognormal_closed(d: Dist, p1: float, p2: float, x_threshold: float, order: int = 96, top_order=None) -> float:
    """Same integral as ParticleDistributions.jl:614-625 with the inner integral in closed form
     and a fixed Gauss-Legendre outer rule in t = ln y.  Used to quantify the
    reference's own adaptive-quadrature error; not the reference algorithm.  ``top_order`` selects the integration range of
    the CUDA rule: [mu - 12 sigma, min(ln x_th, mu + top_order sigma^2 + 12 sigma)]."""
    n, mu, sig = d.n, d.p1, d.p2
    if top_order is None:
        lo = mu + p2 * sig ** 2 - 12 * sig
        hi = math.log(x_threshold)
    else:
        lo = mu - 12 * sig
        hi = min(math.log(x_threshold), mu + top_order * sig ** 2 + 12 * sig)
    if hi <= lo:
        return 0.0
    xs, ws = np.polynomial.legendre.leggauss(order)
    # singular-ish endpoint at y -> x_th (ln(T - y) -> -inf): substitute y = T(1 - s^2)... keep simple:
    t = 0.5 * (hi - lo) * xs + 0.5 * (hi + lo)
    y = np.exp(t)
    fy = n * np.exp(-((t - mu) ** 2) / (2 * sig ** 2)) / (sig * math.sqrt(2 * math.pi))  # f(y) * y
    rem = x_threshold - y
    inner = np.where(rem > 0, n * math.exp(p1 * mu + p1 ** 2 * sig ** 2 / 2) *
                     sp.ndtr((np.log(np.maximum(rem, 1e-300)) - mu - p1 * sig ** 2) / sig), 0.0)
    return float(0.5 * (hi - lo) * np.sum(ws * y ** p2 * fy * inner))


# --------------------------------------------------------------------------------------
# Helper functions — src/helper_functions.jl:13-58
# --------------------------------------------------------------------------------------
def get_dist_moment_ind(NProgMoms: Sequence[int], i: int, m: int) -> int:
    """1-based like the reference (helper_functions.jl:13-21)."""
    if i < 1 or i > len(NProgMoms):
        raise IndexError("distribution index out of range")
    if not (0 < m <= NProgMoms[i - 1]):
        raise ValueError("moment index must be positive integer and equal or smaller than the dist number of prognostic moments!!!")
    return m if i == 1 else sum(NProgMoms[: i - 1]) + m


def get_dist_moments_ind_range(NProgMoms: Sequence[int], i: int) -> range:
    """1-based inclusive range like the reference (helper_functions.jl:29-33)."""
    if i < 1 or i > len(NProgMoms):
        raise IndexError("distribution index out of range")
    last = 0 if i == 1 else sum(NProgMoms[: i - 1])
    return range(last + 1, last + NProgMoms[i - 1] + 1)


def get_moments_normalizing_factors(NProgMoms: Sequence[int], norms: Tuple[float, float]):
    """helper_functions.jl:40-53."""
    if norms[0] <= 0 or norms[1] <= 0:
        raise ValueError("norms must be positive!")
    out = []
    for n_i in NProgMoms:
        for j in range(1, n_i + 1):
            out.append(norms[0] * norms[1] ** (j - 1))
    return out


# --------------------------------------------------------------------------------------
# Kernel tensors — src/Kernels/KernelTensors.jl:44-52, :157-199
# --------------------------------------------------------------------------------------
def check_symmetry(c: np.ndarray):
    if c.size > 1:
        if c.shape[0] != c.shape[1]:
            raise ValueError("array needs to be quadratic in order to be symmetric.")
        for i in range(c.shape[0]):
            for j in range(i + 1, c.shape[0]):
                if c[i, j] != c[j, i]:
                    raise ValueError("array not symmetric.")


def get_normalized_kernel_tensor(c: np.ndarray, norms: Tuple[float, float]) -> np.ndarray:
    """KernelTensors.jl:189-199 (1-based i+j-2 == 0-based a+b)."""
    c = np.asarray(c, dtype=np.float64)
    P = c.shape[0]
    out = np.empty_like(c)
    for a in range(P):
        for b in range(P):
            out[a, b] = c[a, b] * (norms[0] * norms[1] ** float(a + b))
    return out


@dataclass
class CoalescenceData:
    """src/Sources/Coalescence.jl:45-106."""
    N_mom_max: int
    N_2d_ints: Tuple[int, ...]
    dist_thresholds: Tuple[float, ...]
    kernels: list  # [N][N] of (P,P) arrays, normalised
    P: int
    N: int
    moving: bool = False


def make_coalescence_data(kernel, NProgMoms: Sequence[int], dist_thresholds: Sequence[float],
                          norms: Tuple[float, float] = (1.0, 1.0), moving: bool = False) -> CoalescenceData:
    N = len(NProgMoms)
    if isinstance(kernel, np.ndarray) and kernel.ndim == 2:
        kernel = [[kernel for _ in range(N)] for _ in range(N)]
    P = np.asarray(kernel[0][0]).shape[0]
    kernels = [[get_normalized_kernel_tensor(np.asarray(kernel[j][k], dtype=np.float64), norms) for k in range(N)] for j in range(N)]
    for row in kernels:
        for c in row:
            check_symmetry(c)
    N_mom_max = max(NProgMoms) + (P - 1)
    N_2d = []
    for i in range(N):
        if i < N - 1:
            N_2d.append((P - 1) + max(NProgMoms[i], NProgMoms[i + 1]))
        else:
            N_2d.append((P - 1) + NProgMoms[i])
    if not moving:
        thr = tuple(t / norms[1] for t in dist_thresholds)
    else:
        thr = tuple(dist_thresholds)
    return CoalescenceData(N_mom_max, tuple(N_2d), thr, kernels, P, N, moving)


# --------------------------------------------------------------------------------------
# Thresholds — ParticleDistributions.jl:721-761
# --------------------------------------------------------------------------------------
def compute_threshold(d: Dist, percentile: float = 0.97, minx: float = 1e-18) -> float:
    if d.kind == EXPONENTIAL:
        # Julia's log(0.0) is -Inf (percentile 1 → infinite threshold); Python's math.log raises instead
        return max(-d.p1 * (math.log(1 - percentile) if percentile < 1 else -math.inf), minx)
    if d.kind == GAMMA:
        return max(d.p1 * float(sp.gammaincinv(d.p2, percentile)), minx)
    raise TypeError("compute_threshold is defined for Exponential and Gamma only")


def compute_thresholds(pdists: Sequence[Dist], percentiles=0.97):
    N = len(pdists)
    if not isinstance(percentiles, (tuple, list)):
        percentiles = [percentiles] * N
    return tuple(INF if i == N - 1 else compute_threshold(pdists[i], percentiles[i]) for i in range(N))


# --------------------------------------------------------------------------------------
# Coalescence — src/Sources/Coalescence.jl:115-455
# --------------------------------------------------------------------------------------
def get_moments_matrix(pdists: Sequence[Dist], M: int, N_mom_max: int) -> np.ndarray:
    """Coalescence.jl:187-198; result[i, j] for mode i (0-based), order j (0-based)."""
    N = len(pdists)
    out = np.zeros((N, M))
    for j in range(M):
        for i in range(N):
            out[i, j] = moment(pdists[i], float(j)) if (j + 1) <= N_mom_max else 0.0
    return out


def get_finite_2d_integrals(pdists, thresholds, moments, N_2d_ints):
    """Coalescence.jl:200-244."""
    N, M = moments.shape
    out = []
    for i in range(N):
        F = np.zeros((M, M))
        for j in range(1, M + 1):
            for k in range(1, M + 1):
                mm = moments[i, j - 1] * moments[i, k - 1]
                if mm < EPS or k < j or N_2d_ints[i] < j or N_2d_ints[i] < k:
                    v = 0.0
                elif i == N - 1 or math.isinf(thresholds[i]):
                    v = mm
                else:
                    v = jl_min(mm, moment_source_helper(pdists[i], float(j - 1), float(k - 1), thresholds[i]))
                F[j - 1, k - 1] = v
        for j in range(M):
            for k in range(j):
                F[j, k] = F[k, j]
        out.append(F)
    return out


def _binom(n, k):
    return math.comb(n, k)


def get_coal_ints(pdists: Sequence[Dist], cd: CoalescenceData, return_scale: bool = False):
    """get_coal_ints(::AnalyticalCoalStyle, ...) — Coalescence.jl:115-150 (fixed) and :152-185 (moving).

    With ``return_scale`` also returns Σ|terms| per output entry (the cancellation scale used by the
    parity tolerance)."""
    N, P = cd.N, cd.P
    M = P + 2
    NProgMoms = [nparams(d) for d in pdists]
    moments = get_moments_matrix(pdists, M, cd.N_mom_max)
    thresholds = compute_thresholds(pdists, list(cd.dist_thresholds)) if cd.moving else cd.dist_thresholds
    F = get_finite_2d_integrals(pdists, thresholds, moments, cd.N_2d_ints)

    Q = np.zeros((3, N, N))
    R = np.zeros((3, N, N))
    S = np.zeros((3, 2, N))
    Qa = np.zeros((3, N, N))
    Ra = np.zeros((3, N, N))
    Sa = np.zeros((3, 2, N))
    for m in range(3):
        for k in range(N):
            for j in range(N):
                c = cd.kernels[j][k]
                # Q: :260-309
                if not (k <= j or NProgMoms[k] <= m):
                    t = 0.0
                    ta = 0.0
                    for a in range(P):
                        for b in range(P):
                            for cc in range(m + 1):
                                v = c[a, b] * _binom(m, cc) * moments[j, a + cc] * moments[k, b + m - cc]
                                t += v
                                ta += abs(v)
                    Q[m, j, k] = t
                    Qa[m, j, k] = ta
                # R: :311-351
                if not (NProgMoms[k] <= m):
                    t = 0.0
                    ta = 0.0
                    for a in range(P):
                        for b in range(P):
                            v = c[a, b] * moments[j, a] * moments[k, b + m]
                            t += v
                            ta += abs(v)
                    R[m, j, k] = t
                    Ra[m, j, k] = ta
            # S: :353-455
            if k < N - 1 and NProgMoms[k] <= m and NProgMoms[k + 1] <= m:
                continue
            if k == N - 1 and NProgMoms[k] <= m:
                continue
            c = cd.kernels[k][k]
            s1 = s2 = s1a = s2a = 0.0
            for a in range(P):
                for b in range(P):
                    for cc in range(m + 1):
                        f = F[k][a + cc, b + m - cc]
                        v1 = 0.5 * c[a, b] * _binom(m, cc) * f
                        v2 = 0.5 * c[a, b] * _binom(m, cc) * (moments[k, a + cc] * moments[k, b + m - cc] - f)
                        s1 += v1
                        s2 += v2
                        s1a += abs(v1)
                        s2a += abs(0.5 * c[a, b] * _binom(m, cc) * moments[k, a + cc] * moments[k, b + m - cc]) + abs(v1)
            S[m, 0, k], S[m, 1, k] = s1, s2
            Sa[m, 0, k], Sa[m, 1, k] = s1a, s2a

    out = []
    scale = []
    for k in range(N):
        for m in range(NProgMoms[k]):
            v = Q[m, :, k].sum() - R[m, :, k].sum() + S[m, 0, k]
            sc = Qa[m, :, k].sum() + Ra[m, :, k].sum() + Sa[m, 0, k]
            if k > 0:
                v += S[m, 1, k - 1]
                sc += Sa[m, 1, k - 1]
            out.append(v)
            scale.append(sc)
    if return_scale:
        return np.array(out), np.array(scale)
    return np.array(out)


# --------------------------------------------------------------------------------------
# Sedimentation — src/Sources/Sedimentation.jl:22-37
# --------------------------------------------------------------------------------------
def get_sedimentation_flux(pdists: Sequence[Dist], vel: Sequence[Tuple[float, float]]):
    out = []
    for d in pdists:
        for j in range(1, nparams(d) + 1):
            s = 0.0
            for (v, beta) in vel:
                s += -v * moment(d, float(j - 1 + beta))
            out.append(s)
    return np.array(out)


# --------------------------------------------------------------------------------------
# Condensation — src/Sources/Condensation.jl:22-37
# --------------------------------------------------------------------------------------
def get_cond_evap(pdists: Sequence[Dist], s: float, xi: float, rho_l: float = 1000.0):
    out = []
    for d in pdists:
        for j in range(1, nparams(d) + 1):
            if j < 2:
                out.append(0.0)
            else:
                out.append(3 * xi * s * (j - 1) * moment(d, float(j - 1 - 2 / 3)) * (4 * math.pi / 3) ** (2 / 3) / rho_l ** (1 / 3))
    return np.array(out)


def rhs_condensation(mom: Sequence[float], par: "ModelParams", s: float, xi: float):
    """rhs_condensation! — test/examples/utils/box_model_helpers.jl:55-67."""
    pd, _, mom_norms = dists_from_state(mom, par)
    xi_n = xi / par.norms[1] ** (2 / 3)
    return get_cond_evap(pd, s, xi_n) * np.array(mom_norms)


# --------------------------------------------------------------------------------------
# Diagnostics — ParticleDistributions.jl:226-285, :634-687
# --------------------------------------------------------------------------------------
def partial_moment(d: Dist, q: float, x_threshold: float) -> float:
    if d.kind == EXPONENTIAL:
        return d.n * d.p1 ** q * float(sp.gammainc(q + 1.0, x_threshold / d.p1)) * float(sp.gamma(q + 1.0))
    if d.kind == GAMMA:
        return d.n * d.p1 ** q * float(sp.gammainc(q + d.p2, x_threshold / d.p1)) * float(sp.gamma(q + d.p2)) / float(sp.gamma(d.p2))
    if d.kind == MONODISPERSE:
        return 0.0 if x_threshold < d.p1 else d.n * d.p1 ** q
    # Lognormal: quadgk(x -> x^q * dist(x), 0, x_threshold) (:261-269)
    return spint.quad(lambda x: x ** q * float(_lognormal_density_vec(d, x)), 0.0, x_threshold, epsabs=0.0, epsrel=math.sqrt(EPS), limit=200)[0]


def get_standard_N_q(pdists: Sequence[Dist], size_cutoff: float = 1e-6):
    """(N_liq, N_rai, M_liq, M_rai) — ParticleDistributions.jl:634-687."""
    N_liq = sum(partial_moment(d, 0.0, size_cutoff) for d in pdists)
    M_liq = sum(partial_moment(d, 1.0, size_cutoff) for d in pdists)
    N_rai = sum(moment(d, 0.0) - partial_moment(d, 0.0, size_cutoff) for d in pdists)
    M_rai = sum(moment(d, 1.0) - partial_moment(d, 1.0, size_cutoff) for d in pdists)
    return N_liq, N_rai, M_liq, M_rai


# --------------------------------------------------------------------------------------
# RHS glue — test/examples/utils/box_model_helpers.jl:29-53, rainshaft_helpers.jl:45-88
# --------------------------------------------------------------------------------------
@dataclass
class ModelParams:
    kinds: Tuple[int, ...]
    cd: CoalescenceData
    NProgMoms: Tuple[int, ...]
    norms: Tuple[float, float]
    vel: Tuple[Tuple[float, float], ...] = ()
    dz: float = 1.0


def _template(kind: int) -> Dist:
    return Dist(kind, 0.0, 1.0, 1.0)


def dists_from_state(mom: Sequence[float], par: ModelParams):
    mom_norms = get_moments_normalizing_factors(par.NProgMoms, par.norms)
    mn = [m / s for m, s in zip(mom, mom_norms)]
    pd = []
    for i, kind in enumerate(par.kinds):
        rng = get_dist_moments_ind_range(par.NProgMoms, i + 1)
        pd.append(update_dist_from_moments(_template(kind), [mn[r - 1] for r in rng]))
    return pd, mn, mom_norms


def rhs_coal(mom: Sequence[float], par: ModelParams, return_scale: bool = False):
    """box_model_helpers.jl:29-53 (AnalyticalCoalStyle, Fixed or Moving threshold)."""
    pd, _, mom_norms = dists_from_state(mom, par)
    if return_scale:
        ci, sc = get_coal_ints(pd, par.cd, True)
        return ci * np.array(mom_norms), sc * np.array(mom_norms)
    return get_coal_ints(pd, par.cd) * np.array(mom_norms)


def sedimentation_flux_state(mom: Sequence[float], par: ModelParams):
    """rainshaft_helpers.jl:74-77 for one level."""
    pd, _, mom_norms = dists_from_state(mom, par)
    vel_n = tuple((v * par.norms[1] ** b, b) for (v, b) in par.vel)
    return get_sedimentation_flux(pd, vel_n) * np.array(mom_norms)


def rainshaft_rhs(m: np.ndarray, par: ModelParams, return_scale: bool = False):
    """rainshaft_helpers.jl:47-88. ``m`` is (nz, nmom) and is clipped IN PLACE like the reference (:52).
    With ``return_scale`` also returns Σ|terms| per entry (coalescence terms + the two flux terms / dz)."""
    nz, nmom = m.shape
    m[m < 0] = 0
    coal = np.zeros_like(m)
    cscale = np.zeros_like(m)
    flux = np.zeros((nz + 1, nmom))
    for i in range(nz):
        pd, mn, mom_norms = dists_from_state(m[i, :], par)
        if all(v < EPS for v in mn):
            coal[i, :] = 0.0
        else:
            ci, sc = get_coal_ints(pd, par.cd, True)
            coal[i, :] = ci * np.array(mom_norms)
            cscale[i, :] = sc * np.array(mom_norms)
        vel_n = tuple((v * par.norms[1] ** b, b) for (v, b) in par.vel)
        flux[i, :] = get_sedimentation_flux(pd, vel_n) * np.array(mom_norms)
    sed = np.zeros_like(m)
    for i in range(nz):
        sed[i, :] = -(flux[i + 1, :] - flux[i, :]) / par.dz
    if return_scale:
        return coal + sed, cscale + (np.abs(flux[1:, :]) + np.abs(flux[:-1, :])) / par.dz
    return coal + sed


def initial_condition(z: np.ndarray, mom_amp: Sequence[float]) -> np.ndarray:
    """rainshaft_helpers.jl:17-36."""
    zmax = float(np.max(z))
    dz = z[1] - z[0]
    at = np.where((z >= 0.5 * zmax - dz / 2) & (z < 0.75 * zmax - dz / 2), 1.0, 0.0)
    return np.outer(at, np.asarray(mom_amp, dtype=np.float64))


def ssprk33_step(rhs, u, dt):
    """Shu-Osher SSP(3,3) in the operation order of OrdinaryDiffEqSSPRK's SSPRK33 (third-party;
    SURVEY Appendix A.8).  ``rhs`` may clip its argument in place (rainshaft)."""
    k = rhs(u)
    tmp = u + dt * k
    k = rhs(tmp)
    tmp = (3 * u + tmp + dt * k) / 4
    k = rhs(tmp)
    return (u + 2 * tmp + 2 * dt * k) / 3


def ssprk33(rhs, u0, dt, n_steps):
    u = np.array(u0, dtype=np.float64, copy=True)
    for _ in range(n_steps):
        u = ssprk33_step(rhs, u, dt)
    return u


def moment_sums(states: np.ndarray, NProgMoms: Sequence[int]) -> np.ndarray:
    """Σ over parcels of every prognostic moment slot, and Σ over modes per order
    (netcdf_helpers.jl:34-42 analogue).  ``states`` is (n_parcels, Σn_i)."""
    return np.asarray(states, dtype=np.float64).sum(axis=0)
