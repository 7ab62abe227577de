"""Sources/Coalescence — host mirror of src/Sources/Coalescence.jl:45-185 (AnalyticalCoalStyle) and of the
style tags in src/Sources/EquationTypes.jl.  The constructor logic (tensor normalisation, N_mom_max,
N_2d_ints, threshold normalisation) runs on the host exactly as the reference's constructor does; the
tendency evaluation itself is the CUDA kernel."""
import ctypes as C
import itertools
import math
from typing import Sequence, Tuple

import numpy as np

from . import _lib as L
from .context import default_context
from .distributions import nparams
from .kernel_tensors import CoalescenceTensor, get_normalized_kernel_tensor


# ---- EquationTypes.jl:15-22 ---------------------------------------------------------------------------
class AbstractStyle: ...
class CoalescenceStyle(AbstractStyle): ...
class NumericalCoalStyle(CoalescenceStyle): ...
class AnalyticalCoalStyle(CoalescenceStyle): ...
class ThresholdStyle: ...
class MovingThreshold(ThresholdStyle): ...
class FixedThreshold(ThresholdStyle): ...


_uid = itertools.count(1)


def new_uid() -> int:
    """Process-unique token for configuration caching (``id()`` values are recycled after garbage collection)."""
    return next(_uid)


def log_grid(x_threshold: float, n_bins_per_log_unit: int = 15):
    """Node grid of ParticleDistributions.jl:579-582, computed on the host with the host's log10/log
    (SURVEY §7 "parity rule"): (n_bins, x_min, dx)."""
    x_lowerbound = min(1e-5, 1e-5 * x_threshold)
    n_bins = int(math.floor(n_bins_per_log_unit * math.log10(x_threshold / x_lowerbound)))
    x_min = math.log(x_lowerbound)
    dx = (math.log(x_threshold) - math.log(x_lowerbound)) / n_bins
    return n_bins, x_min, dx


class CoalescenceData:
    """CoalescenceData{N,P,FT} — Coalescence.jl:45-106.

    ``kernel``: one CoalescenceTensor for every pair, or an N×N nested sequence of tensors."""

    def __init__(self, kernel, NProgMoms: Sequence[int], dist_thresholds: Sequence[float],
                 norms: Tuple[float, float] = (1.0, 1.0), ts: ThresholdStyle = None):
        ts = ts if ts is not None else FixedThreshold()
        N = len(NProgMoms)
        if isinstance(kernel, CoalescenceTensor):
            kernel = tuple(tuple(kernel for _ in range(N)) for _ in range(N))
        if len(kernel) != N or any(len(row) != N for row in kernel):
            raise ValueError("kernel matrix must be N x N")
        if len(dist_thresholds) != N:
            raise ValueError("one threshold per distribution is required")
        P = kernel[0][0].P
        if any(k.P != P for row in kernel for k in row):
            raise ValueError("all kernel tensors must have the same order")
        self.kernels = tuple(tuple(get_normalized_kernel_tensor(kernel[j][k], norms) for k in range(N)) for j in range(N))
        self.N, self.P = N, P
        self.NProgMoms = tuple(int(v) for v in NProgMoms)
        self.N_mom_max = max(NProgMoms) + (P - 1)
        self.N_2d_ints = tuple((P - 1) + (max(NProgMoms[i], NProgMoms[i + 1]) if i < N - 1 else NProgMoms[i])
                               for i in range(N))
        self.threshold_style = ts
        if isinstance(ts, FixedThreshold):
            self.dist_thresholds = tuple(float(t) / norms[1] for t in dist_thresholds)
        else:
            self.dist_thresholds = tuple(float(t) for t in dist_thresholds)
        self.norms = (float(norms[0]), float(norms[1]))
        self.uid = new_uid()


def build_config(kinds: Sequence[int], coal_data: CoalescenceData, norms=None, vel=(), dz: float = 1.0, nz: int = 1,
                 k_range=None, moving=None) -> L.cloudy_config:
    """Marshal CoalescenceData + the drivers' ODE_parameters into the C-ABI struct."""
    cd = coal_data
    N, P = cd.N, cd.P
    if N > L.MAX_MODES or P > L.MAX_P:
        raise ValueError("configuration exceeds CLOUDY_MAX_MODES / CLOUDY_MAX_P")
    if len(kinds) != N:
        raise ValueError("one distribution kind per mode is required")
    cfg = L.cloudy_config()
    cfg.n_modes, cfg.P = N, P
    for i in range(N):
        cfg.kind[i] = int(kinds[i])
        cfg.nprog[i] = int(cd.NProgMoms[i])
        cfg.n_2d_ints[i] = int(cd.N_2d_ints[i])
        cfg.thresholds[i] = cd.dist_thresholds[i]
    if moving is None:
        moving = isinstance(cd.threshold_style, MovingThreshold)
    cfg.threshold_style = L.MOVING_THRESHOLD if moving else L.FIXED_THRESHOLD
    cfg.n_mom_max = cd.N_mom_max
    cfg.bins_per_log_unit = 15
    for j in range(N):
        for k in range(N):
            for a in range(P):
                for b in range(P):
                    cfg.c[j][k][a][b] = float(cd.kernels[j][k].c[a, b])
    if cfg.threshold_style == L.FIXED_THRESHOLD:
        for i in range(N - 1):
            t = cd.dist_thresholds[i]
            if math.isfinite(t) and kinds[i] in (L.GAMMA, L.EXPONENTIAL):
                if not t > 0:
                    raise ValueError("thresholds must be positive")
                nb, x_min, dx = log_grid(t)
                cfg.n_bins[i], cfg.x_min[i], cfg.dx[i] = nb, x_min, dx
    nrm = norms if norms is not None else cd.norms
    cfg.norms[0], cfg.norms[1] = float(nrm[0]), float(nrm[1])
    kr = k_range if k_range is not None else (float(np.finfo(np.float64).eps), 10.0)
    cfg.k_range[0], cfg.k_range[1] = kr
    if len(vel) > L.MAX_VEL:
        raise ValueError("too many terminal-velocity terms")
    cfg.n_vel = len(vel)
    for i, (v, b) in enumerate(vel):
        cfg.vel[i][0], cfg.vel[i][1] = float(v), float(b)
    cfg.dz = float(dz)
    cfg.nz = int(nz)
    return cfg


def apply_config(ctx, cfg: L.cloudy_config, key=None):
    if key is not None and ctx.config == key:
        return
    L.check(L.load().cloudy_config_set(ctx.handle, C.byref(cfg)))
    ctx.config = key


def get_coal_ints(cs, pdists, coal_data: CoalescenceData, ts: ThresholdStyle = None, ctx=None):
    """get_coal_ints(::AnalyticalCoalStyle, pdists, coal_data[, ::MovingThreshold]) — Coalescence.jl:115-185.
    Returns the flat tuple of coalescence integrals (normalised units), one per prognostic moment."""
    if not isinstance(cs, AnalyticalCoalStyle):
        raise NotImplementedError("only AnalyticalCoalStyle is accelerated (NumericalCoalStyle is out of scope, SURVEY §2 #2)")
    if len(pdists) != coal_data.N:
        raise ValueError("number of distributions does not match coal_data")
    for d, n in zip(pdists, coal_data.NProgMoms):
        if nparams(d) != n:
            raise ValueError("NProgMoms does not match the distributions")
    ctx = ctx or default_context()
    kinds = tuple(d.kind for d in pdists)
    # Only the 4-argument method (Coalescence.jl:152-157) reads coal_data.dist_thresholds as percentiles; the 3-argument method
    # (:115-150) ALWAYS treats them as mass thresholds, whatever style coal_data was built with (the reference's
    # CoalescenceData does not even store the style).
    moving = isinstance(ts, MovingThreshold)
    cfg = build_config(kinds, coal_data, norms=(1.0, 1.0), moving=moving)
    apply_config(ctx, cfg, key=("coal_ints", coal_data.uid, kinds, moving))
    params = np.zeros((coal_data.N, 3))
    for i, d in enumerate(pdists):
        p = d.params()
        params[i, : len(p)] = p
    out = np.zeros(sum(coal_data.NProgMoms))
    L.check(L.load().cloudy_get_coal_ints_1(ctx.handle, L.dptr(params), L.dptr(out)))
    return tuple(out.tolist())
