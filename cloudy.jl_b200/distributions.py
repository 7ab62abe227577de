"""ParticleDistributions — host mirror of src/ParticleDistributions/ParticleDistributions.jl.

The structs, constructors' validation and names follow the reference; every numerical method is
evaluated on the GPU through the C ABI (no host arithmetic path)."""
import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .context import default_context

EPS = float(np.finfo(np.float64).eps)


class PrimitiveParticleDistribution:
    """abstract — ParticleDistributions.jl:51"""
    kind = -1

    def params(self):
        raise NotImplementedError

    def __call__(self, x):
        return density(self, x)


@dataclass(frozen=True)
class ExponentialPrimitiveParticleDistribution(PrimitiveParticleDistribution):
    """ParticleDistributions.jl:66-79"""
    n: float
    θ: float
    kind = L.EXPONENTIAL

    def __post_init__(self):
        if self.n < 0 or self.θ <= 0:
            raise ValueError("n needs to be nonnegative. θ needs to be positive.")

    def params(self):
        return (float(self.n), float(self.θ))


@dataclass(frozen=True)
class GammaPrimitiveParticleDistribution(PrimitiveParticleDistribution):
    """ParticleDistributions.jl:93-107"""
    n: float
    θ: float
    k: float
    kind = L.GAMMA

    def __post_init__(self):
        if self.n < 0 or self.θ <= 0 or self.k <= 0:
            raise ValueError("n needs to be nonnegative. θ and k need to be positive.")

    def params(self):
        return (float(self.n), float(self.θ), float(self.k))


@dataclass(frozen=True)
class MonodispersePrimitiveParticleDistribution(PrimitiveParticleDistribution):
    """ParticleDistributions.jl:120-132"""
    n: float
    θ: float
    kind = L.MONODISPERSE

    def __post_init__(self):
        if self.n < 0 or self.θ <= 0:
            raise ValueError("n needs to be nonnegative. θ needs to be positive.")

    def params(self):
        return (float(self.n), float(self.θ))


@dataclass(frozen=True)
class LognormalPrimitiveParticleDistribution(PrimitiveParticleDistribution):
    """ParticleDistributions.jl:145-159"""
    n: float
    μ: float
    σ: float
    kind = L.LOGNORMAL

    def __post_init__(self):
        if self.n < 0 or self.σ <= 0:
            raise ValueError("n needs to be nonnegative. σ needs to be positive.")

    def params(self):
        return (float(self.n), float(self.μ), float(self.σ))


_CTORS = {
    L.EXPONENTIAL: ExponentialPrimitiveParticleDistribution,
    L.GAMMA: GammaPrimitiveParticleDistribution,
    L.LOGNORMAL: LognormalPrimitiveParticleDistribution,
    L.MONODISPERSE: MonodispersePrimitiveParticleDistribution,
}


def nparams(dist) -> int:
    """ParticleDistributions.jl:425-427"""
    return len(dist.params())


def _params3(dist):
    p = list(dist.params()) + [1.0]
    return (C.c_double * 3)(*p[:3])


def moment(dist, q: float, ctx=None) -> float:
    """moment(dist, q) — ParticleDistributions.jl:216"""
    ctx = ctx or default_context()
    out = C.c_double()
    L.check(L.load().cloudy_moment(ctx.handle, dist.kind, _params3(dist), float(q), C.byref(out)))
    return out.value


def get_moments(dist, ctx=None):
    """ParticleDistributions.jl:293-315 — the first nparams(dist) integer moments."""
    return [moment(dist, float(q), ctx) for q in range(nparams(dist))]


def update_dist_from_moments(dist, moments, param_range=None, ctx=None):
    """ParticleDistributions.jl:456-541.  ``param_range``: {"k": (lo, hi)} or {"μ": (lo, hi), "σ": (lo, hi)}."""
    if len(moments) != nparams(dist):
        raise TypeError("MethodError: wrong number of moments for this distribution")
    ctx = ctx or default_context()
    rng = None
    if param_range is not None:
        if dist.kind == L.GAMMA:
            rng = (C.c_double * 4)(param_range["k"][0], param_range["k"][1], 0.0, 0.0)
        elif dist.kind == L.LOGNORMAL:
            mu = param_range.get("μ", (-math.inf, math.inf))
            sg = param_range.get("σ", (EPS, math.inf))
            rng = (C.c_double * 4)(mu[0], mu[1], sg[0], sg[1])
    m = (C.c_double * 3)(*(list(map(float, moments)) + [0.0])[:3])
    out = (C.c_double * 3)()
    inv = C.c_int32()
    L.check(L.load().cloudy_update_dist_from_moments(ctx.handle, dist.kind, m, rng, out, C.byref(inv)))
    if inv.value:
        raise ValueError("DomainError: sqrt of a negative number (ParticleDistributions.jl:498)")
    return _CTORS[dist.kind](*list(out)[: nparams(dist)])


def moment_source_helper(dist, p1: float, p2: float, x_threshold: float, n_bins_per_log_unit: int = 15, ctx=None) -> float:
    """ParticleDistributions.jl:557-625"""
    ctx = ctx or default_context()
    out = C.c_double()
    L.check(L.load().cloudy_moment_source_helper(ctx.handle, dist.kind, _params3(dist), float(p1), float(p2),
                                                 float(x_threshold), int(n_bins_per_log_unit), C.byref(out)))
    return out.value


def integrate_SimpsonEvenFast(n_bins: int, dx: float, y, ctx=None) -> float:
    """ParticleDistributions.jl:698-710; ``y(j)`` 1-based callable as in the reference."""
    if n_bins < 3:
        raise ValueError("n_bins must be at least 3")
    ctx = ctx or default_context()
    tab = np.array([y(j) for j in range(1, n_bins + 2)], dtype=np.float64)
    out = C.c_double()
    L.check(L.load().cloudy_integrate_simpson(ctx.handle, int(n_bins), float(dx), L.dptr(tab), C.byref(out)))
    return out.value


def compute_threshold(pdist, percentile: float = 0.97, minx: float = 1e-18, ctx=None) -> float:
    """compute_threshold(pdist, percentile, minx) — ParticleDistributions.jl:747-761 (Exponential and Gamma only)."""
    if pdist.kind not in (L.EXPONENTIAL, L.GAMMA):
        raise TypeError("MethodError: no method matching compute_threshold for this distribution")
    ctx = ctx or default_context()
    out = C.c_double()
    L.check(L.load().cloudy_compute_threshold(ctx.handle, pdist.kind, _params3(pdist), float(percentile), float(minx), C.byref(out)))
    return out.value


def compute_thresholds(pdists, percentile=0.97, ctx=None):
    """compute_thresholds(pdists, percentile | percentiles) — ParticleDistributions.jl:721-745: Inf for the last mode."""
    N = len(pdists)
    pcs = list(percentile) if isinstance(percentile, (tuple, list)) else [percentile] * N
    return tuple(math.inf if i == N - 1 else compute_threshold(pdists[i], pcs[i], ctx=ctx) for i in range(N))


def normed_density(dist, x: float) -> float:
    """ParticleDistributions.jl:411-416 — plotting helper, host arithmetic."""
    if x < 0:
        raise ValueError("Density can only be evaluated at nonnegative values.")
    if dist.kind == L.MONODISPERSE:
        raise TypeError("MethodError: normed_density_func is not defined for Monodisperse distributions")
    return density(type(dist)(1.0, *dist.params()[1:]), x)


def density(dist, x: float) -> float:
    """ParticleDistributions.jl:397-402 — not on the accelerated path (plotting helper); host arithmetic."""
    if x < 0:
        raise ValueError("Density can only be evaluated at nonnegative values.")
    if dist.kind == L.EXPONENTIAL:
        return dist.n / dist.θ * math.exp(-x / dist.θ)
    if dist.kind == L.GAMMA:
        return dist.n * x ** (dist.k - 1) / dist.θ ** dist.k / math.gamma(dist.k) * math.exp(-x / dist.θ)
    if dist.kind == L.LOGNORMAL:
        if x == 0:
            return float("nan")
        return dist.n * math.exp(-((math.log(x) - dist.μ) ** 2 / (2 * dist.σ ** 2))) / (x * dist.σ * math.sqrt(2 * math.pi))
    return dist.n / (2 * dist.θ / 10.0) if abs(x - dist.θ) < dist.θ / 10.0 else 0.0
