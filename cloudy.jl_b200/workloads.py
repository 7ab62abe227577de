"""Synthetic workloads of BASELINE.json's configs: run-constant model parameters plus a seeded
generator of parcel / cell states.  Pure numpy host code; used by tests and bench.py."""
import math
from types import SimpleNamespace

import numpy as np

from . import _lib as L
from .coalescence import CoalescenceData
from .distributions import (ExponentialPrimitiveParticleDistribution as Exp, GammaPrimitiveParticleDistribution as Gam,
                            LognormalPrimitiveParticleDistribution as LogN, MonodispersePrimitiveParticleDistribution as Mono)
from .ensemble import ModelParameters
from .kernel_tensors import CoalescenceTensor, HydrodynamicKernelFunction, LongKernelFunction, polyfit

SEED0 = 20250117
NORMS = (1e6, 1e-9)


def _norm_factors(NProgMoms, norms):
    return np.array([norms[0] * norms[1] ** q for n in NProgMoms for q in range(n)])


def _moments_from_params(kind, n, a, b, nprog):
    """Normalised moments M_0..M_{nprog-1} (vectorised get_moments, ParticleDistributions.jl:293-315)."""
    if kind == L.GAMMA:
        cols = [n, n * b * a, n * b * (b + 1) * a ** 2]
    elif kind == L.LOGNORMAL:
        cols = [n, n * np.exp(a + b ** 2 / 2), n * np.exp(2.0 * a + 2.0 * b ** 2)]
    else:
        cols = [n, n * a]
    return np.stack(cols[:nprog], axis=1)


def _logu(rng, lo, hi, size):
    return np.exp(rng.uniform(math.log(lo), math.log(hi), size))


def linear_tensor(B=5.0):
    """Exact tensor of LinearKernelFunction(B): K = B (x + y) (box_gamma_mixture.jl:20-21)."""
    return CoalescenceTensor(np.array([[0.0, B], [B, 0.0]]))


def c1_smoluchowski():
    """C1: single Exponential mode, constant kernel, 1 parcel (test_Sources_correctness.jl:41-64)."""
    ker = CoalescenceTensor(np.array([[1.0]]))
    pd = (Exp(1.0, 1.0),)
    par = ModelParameters(pdists=pd, coal_data=CoalescenceData(ker, (2,), (math.inf,)), NProgMoms=(2,), norms=(1.0, 1.0), dt=1e-4)
    return par, np.array([[1.0, 2.0]])


def c2_gamma_exp(n_parcels=1 << 20, seed=SEED0 + 2, empty_frac=0.02):
    """C2/C5: Gamma cloud + Exponential rain, Golovin kernel, thresholds (0.5, Inf) normalised."""
    rng = np.random.default_rng(seed)
    NProgMoms = (3, 2)
    pd = (Gam(1e8, 1e-10, 1.0), Exp(1.0, 1e-8))
    cd = CoalescenceData(linear_tensor(5.0), NProgMoms, (5e-10, math.inf), NORMS)
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=NORMS, dt=10.0)
    n1 = _logu(rng, 1e1, 1e3, n_parcels); th1 = _logu(rng, 0.03, 0.3, n_parcels); k1 = _logu(rng, 0.5, 5.0, n_parcels)
    n2 = _logu(rng, 1e-6, 1e0, n_parcels); th2 = _logu(rng, 1.0, 30.0, n_parcels)
    m = np.concatenate([_moments_from_params(L.GAMMA, n1, th1, k1, 3), _moments_from_params(L.EXPONENTIAL, n2, th2, None, 2)], axis=1)
    empty = rng.random(n_parcels) < empty_frac
    m[empty, :] = 0.0
    return par, m * _norm_factors(NProgMoms, NORMS)


def c2_gamma_gamma(n_parcels=4096, seed=SEED0 + 22):
    """The 6-moment Gamma+Gamma variant of box_gamma_mixture.jl:14-28."""
    rng = np.random.default_rng(seed)
    NProgMoms = (3, 3)
    pd = (Gam(1e8, 1e-10, 1.0), Gam(1.0, 1e-8, 1.0))
    cd = CoalescenceData(linear_tensor(5.0), NProgMoms, (5e-10, math.inf), NORMS)
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=NORMS, dt=10.0)
    n1 = _logu(rng, 1e1, 1e3, n_parcels); th1 = _logu(rng, 0.03, 0.3, n_parcels); k1 = _logu(rng, 0.5, 5.0, n_parcels)
    n2 = _logu(rng, 1e-6, 1e0, n_parcels); th2 = _logu(rng, 1.0, 30.0, n_parcels); k2 = _logu(rng, 0.5, 5.0, n_parcels)
    m = np.concatenate([_moments_from_params(L.GAMMA, n1, th1, k1, 3), _moments_from_params(L.GAMMA, n2, th2, k2, 3)], axis=1)
    m[0] = np.array([1e8, 1e-2, 2e-12, 1, 1e-8, 2e-16]) / _norm_factors(NProgMoms, NORMS)  # the script's own IC
    return par, m * _norm_factors(NProgMoms, NORMS)


def c3_rainshaft(n_columns=4096, nz=256, seed=SEED0 + 3):
    """C3: rainshaft, 2 Gamma modes, thresholds (0.2, Inf) normalised, vel = 50 x^(1/6) (rainshaft_gamma_mixture.jl:13-49)."""
    rng = np.random.default_rng(seed)
    NProgMoms = (3, 3)
    pd = (Gam(1e7, 1e-10, 1.0), Gam(0.0, 1e-9, 1.0))
    cd = CoalescenceData(linear_tensor(5.0), NProgMoms, (2e-10, math.inf), NORMS)
    zmax = 3000.0
    dz = zmax / nz
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=NORMS, vel=((50.0, 1.0 / 6),), dz=dz, dt=1.0)
    z = np.arange(dz / 2, zmax, dz)[:nz]
    at = np.where((z >= 0.5 * zmax - dz / 2) & (z < 0.75 * zmax - dz / 2), 1.0, 0.0)  # rainshaft_helpers.jl:21-27
    amp = np.array([1e7, 1e-3, 2e-13, 0.0, 0.0, 0.0])
    fac = _logu(rng, 0.3, 3.0, n_columns)
    m = fac[:, None, None] * at[None, :, None] * amp[None, None, :]
    return par, m  # (n_columns, nz, 6)


def hydro_tensor(order=4):
    """polyfit of HydrodynamicKernelFunction(1e2 π) on the reference's grid (box_gamma_mixture_hydro.jl:22-23)."""
    return CoalescenceTensor(polyfit(HydrodynamicKernelFunction(1e2 * math.pi), order, 1e-6, 0.0, NORMS))


def c4_three_modes(n_parcels=1 << 24, seed=SEED0 + 4, last="gamma", empty_frac=0.02):
    """C4: 3 modes (Gamma, Gamma, Gamma|Lognormal), order-4 hydrodynamic tensor, thresholds (1, 100, Inf) normalised
    (box_gamma_mixture_3modes.jl:29, box_gamma_mixture_hydro.jl:22-23)."""
    rng = np.random.default_rng(seed)
    NProgMoms = (3, 3, 3)
    last_kind = L.LOGNORMAL if last == "lognormal" else L.GAMMA
    pd = (Gam(1e8, 1e-10, 1.0), Gam(0.0, 1e-8, 1.0), LogN(0.0, 1.0, 1.0) if last_kind == L.LOGNORMAL else Gam(0.0, 1e-6, 1.0))
    cd = CoalescenceData(hydro_tensor(4), NProgMoms, (1e-9, 1e-7, math.inf), NORMS)
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=NORMS, dt=1.0)
    n1 = _logu(rng, 1e1, 1e3, n_parcels); th1 = _logu(rng, 0.03, 0.5, n_parcels); k1 = _logu(rng, 0.5, 5.0, n_parcels)
    n2 = _logu(rng, 1e-4, 1e1, n_parcels); th2 = _logu(rng, 2.0, 40.0, n_parcels); k2 = _logu(rng, 0.5, 5.0, n_parcels)
    n3 = _logu(rng, 1e-8, 1e-3, n_parcels)
    if last_kind == L.LOGNORMAL:
        mu3 = rng.uniform(math.log(100.0), math.log(1000.0), n_parcels); s3 = rng.uniform(0.3, 0.9, n_parcels)
        m3 = _moments_from_params(L.LOGNORMAL, n3, mu3, s3, 3)
    else:
        th3 = _logu(rng, 100.0, 1000.0, n_parcels); k3 = _logu(rng, 0.5, 5.0, n_parcels)
        m3 = _moments_from_params(L.GAMMA, n3, th3, k3, 3)
    m = np.concatenate([_moments_from_params(L.GAMMA, n1, th1, k1, 3), _moments_from_params(L.GAMMA, n2, th2, k2, 3), m3], axis=1)
    empty = rng.random(n_parcels) < empty_frac
    m[empty, :] = 0.0
    return par, m * _norm_factors(NProgMoms, NORMS)


def long_kernel_two_modes(n_parcels=1024, seed=SEED0 + 44):
    """Long-type variant: a matrix of P = 3 tensors (box_gamma_mixture_long.jl:21-36)."""
    rng = np.random.default_rng(seed)
    NProgMoms = (3, 3)
    kf = LongKernelFunction(5.236e-10, 9.44e9, 5.78)
    t11 = CoalescenceTensor(kf, 2, 5e-10)
    tot = CoalescenceTensor(kf, 2, 1e-6, 5e-10)
    kern = ((t11, tot), (tot, tot))
    pd = (Gam(1e7, 1e-10, 1.0), Gam(1e5, 1e-9, 1.0))
    cd = CoalescenceData(kern, NProgMoms, (5e-10, math.inf), NORMS)
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=NORMS, dt=1.0)
    n1 = _logu(rng, 1e0, 1e2, n_parcels); th1 = _logu(rng, 0.03, 0.3, n_parcels); k1 = _logu(rng, 0.5, 5.0, n_parcels)
    n2 = _logu(rng, 1e-3, 1e0, n_parcels); th2 = _logu(rng, 0.5, 5.0, n_parcels); k2 = _logu(rng, 0.5, 5.0, n_parcels)
    m = np.concatenate([_moments_from_params(L.GAMMA, n1, th1, k1, 3), _moments_from_params(L.GAMMA, n2, th2, k2, 3)], axis=1)
    m[0] = np.array([1e7, 1e-3, 2e-13, 1e5, 1e-4, 2e-13]) / _norm_factors(NProgMoms, NORMS)
    return par, m * _norm_factors(NProgMoms, NORMS)


def mono_gamma(n_parcels=1024, seed=SEED0 + 45):
    """Monodisperse + Gamma mixture (box_mono_gamma_mixture.jl:14-28)."""
    rng = np.random.default_rng(seed)
    NProgMoms = (2, 3)
    pd = (Mono(1e7, 1e-10), Gam(1e5, 1e-9, 1.0))
    cd = CoalescenceData(linear_tensor(5.0), NProgMoms, (5e-10, math.inf), NORMS)
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=NORMS, dt=1.0)
    n1 = _logu(rng, 1e0, 1e2, n_parcels); th1 = _logu(rng, 0.05, 0.6, n_parcels)
    n2 = _logu(rng, 1e-3, 1e0, n_parcels); th2 = _logu(rng, 0.5, 5.0, n_parcels); k2 = _logu(rng, 0.5, 5.0, n_parcels)
    m = np.concatenate([_moments_from_params(L.MONODISPERSE, n1, th1, None, 2), _moments_from_params(L.GAMMA, n2, th2, k2, 3)], axis=1)
    m[0] = np.array([1e7, 1e-3, 1e5, 1e-4, 2e-13]) / _norm_factors(NProgMoms, NORMS)
    return par, m * _norm_factors(NProgMoms, NORMS)


def moving_four_modes(n_parcels=512, seed=SEED0 + 46):
    """MovingThreshold, 4 Gamma modes, percentiles (0.99, 0.99, 0.99, 1.0) (box_gamma_mix_moving.jl:14-44)."""
    from .coalescence import MovingThreshold
    rng = np.random.default_rng(seed)
    NProgMoms = (3, 3, 3, 3)
    pd = (Gam(1e8, 1e-10, 1.0), Gam(0.0, 1e-8, 1.0), Gam(0.0, 1e-6, 1.0), Gam(0.0, 1e-4, 1.0))
    cd = CoalescenceData(linear_tensor(5.0), NProgMoms, (0.99, 0.99, 0.99, 1.0), NORMS, MovingThreshold())
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=NORMS, dt=1.0)
    cols = []
    for (nlo, nhi, tlo, thi) in ((1e1, 1e3, 0.03, 0.3), (1e-3, 1e0, 3.0, 30.0), (1e-6, 1e-3, 3e2, 3e3), (1e-9, 1e-6, 3e4, 3e5)):
        n = _logu(rng, nlo, nhi, n_parcels); th = _logu(rng, tlo, thi, n_parcels); k = _logu(rng, 0.5, 5.0, n_parcels)
        cols.append(_moments_from_params(L.GAMMA, n, th, k, 3))
    m = np.concatenate(cols, axis=1)
    m[0] = np.array([1e8, 1e-2, 2e-12, 0, 0, 0, 0, 0, 0, 0, 0, 0]) / _norm_factors(NProgMoms, NORMS)  # the script's own IC
    m[1, 6:] = 0.0   # two trailing modes empty
    return par, m * _norm_factors(NProgMoms, NORMS)


def moving_gamma_exp(n_parcels=512, seed=SEED0 + 47, percentile=0.97):
    """MovingThreshold with a Gamma cloud and an Exponential rain mode (default percentile of compute_thresholds)."""
    from .coalescence import MovingThreshold
    par, m = c2_gamma_exp(n_parcels, seed=seed)
    cd = CoalescenceData(linear_tensor(5.0), par.NProgMoms, (percentile, 1.0), NORMS, MovingThreshold())
    par2 = ModelParameters(pdists=par.pdists, coal_data=cd, NProgMoms=par.NProgMoms, norms=NORMS, dt=1.0)
    return par2, m


def lognormal_mixture(n_parcels=256, seed=SEED0 + 48):
    """Two Lognormal modes, the first with a finite threshold (box_lognorm_mixture.jl:14-28)."""
    rng = np.random.default_rng(seed)
    NProgMoms = (3, 3)
    pd = (LogN(1e7, -23.37, 0.833), LogN(1e5, -21.07, 0.833))
    cd = CoalescenceData(linear_tensor(5.0), NProgMoms, (5e-10, math.inf), NORMS)
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=NORMS, dt=1.0)
    n1 = _logu(rng, 1e0, 1e2, n_parcels); mu1 = rng.uniform(math.log(0.03), math.log(0.4), n_parcels); s1 = rng.uniform(0.3, 1.0, n_parcels)
    n2 = _logu(rng, 1e-3, 1e0, n_parcels); mu2 = rng.uniform(math.log(0.5), math.log(5.0), n_parcels); s2 = rng.uniform(0.3, 1.0, n_parcels)
    m = np.concatenate([_moments_from_params(L.LOGNORMAL, n1, mu1, s1, 3), _moments_from_params(L.LOGNORMAL, n2, mu2, s2, 3)], axis=1)
    m[0] = np.array([1e7, 1e-3, 2e-13, 1e5, 1e-4, 2e-13]) / _norm_factors(NProgMoms, NORMS)
    return par, m * _norm_factors(NProgMoms, NORMS)


def three_modes_order2(n_parcels=256, seed=SEED0 + 49):
    """Three modes with an order-2 tensor (N = 3, P = 3), a shape none of the reference's drivers uses.
    Exponential + Gamma + Gamma modes, symmetric order-2 tensor, thresholds (1, 50, Inf) normalised."""
    rng = np.random.default_rng(seed)
    NProgMoms = (2, 3, 3)
    pd = (Exp(1e8, 1e-10), Gam(0.0, 1e-8, 1.0), Gam(0.0, 1e-6, 1.0))
    c = np.array([[1e-9, 4.0, 2e8], [4.0, 3e8, 1e16], [2e8, 1e16, 0.0]])  # un-normalised: c_ab / (1e6 * 1e-9^(a+b)) scale
    cd = CoalescenceData(CoalescenceTensor(c), NProgMoms, (1e-9, 5e-8, math.inf), NORMS)
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=NORMS, dt=1.0)
    n1 = _logu(rng, 1e1, 1e3, n_parcels); th1 = _logu(rng, 0.03, 0.5, n_parcels)
    n2 = _logu(rng, 1e-4, 1e1, n_parcels); th2 = _logu(rng, 2.0, 40.0, n_parcels); k2 = _logu(rng, 0.5, 5.0, n_parcels)
    n3 = _logu(rng, 1e-8, 1e-3, n_parcels); th3 = _logu(rng, 100.0, 1000.0, n_parcels); k3 = _logu(rng, 0.5, 5.0, n_parcels)
    m = np.concatenate([_moments_from_params(L.EXPONENTIAL, n1, th1, None, 2), _moments_from_params(L.GAMMA, n2, th2, k2, 3),
                        _moments_from_params(L.GAMMA, n3, th3, k3, 3)], axis=1)
    return par, m * _norm_factors(NProgMoms, NORMS)


def random_model(rng, N, P, kinds=None, n_parcels=64, finite_thresholds=True, moving=False):
    """A random but physically shaped configuration for fuzz tests: ``kinds`` per mode (default: random Exp/Gamma, last mode
    any of the four), symmetric random tensor with the magnitude pattern of a normalised kernel, thresholds increasing by mode."""
    if kinds is None:
        # compute_threshold exists for Exponential and Gamma modes only (ParticleDistributions.jl:747-761)
        inner = [L.EXPONENTIAL, L.GAMMA] if moving else [L.EXPONENTIAL, L.GAMMA, L.MONODISPERSE]
        kinds = [int(rng.choice(inner)) for _ in range(N - 1)] + \
                [int(rng.choice([L.EXPONENTIAL, L.GAMMA, L.MONODISPERSE, L.LOGNORMAL]))]
    ctor = {L.EXPONENTIAL: lambda: Exp(1.0, 1.0), L.GAMMA: lambda: Gam(1.0, 1.0, 1.0), L.MONODISPERSE: lambda: Mono(1.0, 1.0),
            L.LOGNORMAL: lambda: LogN(1.0, 0.0, 1.0)}
    pd = tuple(ctor[k]() for k in kinds)
    NProgMoms = tuple(3 if k in (L.GAMMA, L.LOGNORMAL) else 2 for k in kinds)
    c = np.zeros((P, P))
    for a in range(P):
        for b in range(a, P):
            c[a, b] = c[b, a] = rng.uniform(0.2, 1.0) * 5e-3 * 10.0 ** (-2.5 * (a + b - 1)) if (a + b) > 0 else rng.uniform(0.0, 1e-3)
    scales = [0.1 * 30.0 ** i for i in range(N)]  # mean mass scale of mode i (normalised units)
    thr = tuple((5.0 * scales[i] if finite_thresholds and rng.random() < 0.85 else math.inf) if i < N - 1 else math.inf for i in range(N))
    if moving:
        from .coalescence import MovingThreshold
        pct = tuple((float(rng.choice([0.5, 0.9, 0.97, 0.99])) if rng.random() < 0.85 else 1.0) if i < N - 1 else 1.0 for i in range(N))
        cd = CoalescenceData(CoalescenceTensor(c), NProgMoms, pct, (1.0, 1.0), MovingThreshold())
    else:
        cd = CoalescenceData(CoalescenceTensor(c), NProgMoms, thr, (1.0, 1.0))
    par = ModelParameters(pdists=pd, coal_data=cd, NProgMoms=NProgMoms, norms=(1.0, 1.0), dt=1.0)
    cols = []
    for i, k in enumerate(kinds):
        n = _logu(rng, 1e-2, 1e2, n_parcels) * 100.0 ** (-i)
        th = _logu(rng, 0.3, 3.0, n_parcels) * scales[i]
        if k == L.GAMMA:
            cols.append(_moments_from_params(k, n, th, _logu(rng, 0.4, 6.0, n_parcels), 3))
        elif k == L.LOGNORMAL:
            cols.append(_moments_from_params(k, n, np.log(th), rng.uniform(0.3, 0.9, n_parcels), 3))
        else:
            cols.append(_moments_from_params(k, n, th, None, 2))
    m = np.concatenate(cols, axis=1)
    m[rng.random(n_parcels) < 0.05] = 0.0
    return par, m
