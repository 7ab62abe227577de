"""Sources/Sedimentation — host mirror of src/Sources/Sedimentation.jl:22-37."""
import ctypes as C

import numpy as np

from . import _lib as L
from .context import default_context
from .distributions import nparams


def get_sedimentation_flux(pdists, vel, ctx=None):
    """Sedimentation flux of all prognostic moments for terminal velocity Σ vel[i][0] * x^vel[i][1]."""
    ctx = ctx or default_context()
    N = len(pdists)
    kinds = (C.c_int32 * N)(*[d.kind for d in pdists])
    params = np.zeros((N, 3))
    for i, d in enumerate(pdists):
        p = d.params()
        params[i, : len(p)] = p
        if len(p) < 3:
            params[i, 2] = 1.0
    v = np.array([[float(a), float(b)] for a, b in vel], dtype=np.float64).reshape(-1)
    out = np.zeros(sum(nparams(d) for d in pdists))
    L.check(L.load().cloudy_get_sedimentation_flux_1(ctx.handle, N, kinds, L.dptr(params), len(vel), L.dptr(v), L.dptr(out)))
    return tuple(out.tolist())


def _pack(pdists):
    N = len(pdists)
    kinds = (C.c_int32 * N)(*[d.kind for d in pdists])
    params = np.zeros((N, 3))
    for i, d in enumerate(pdists):
        p = d.params()
        params[i, : len(p)] = p
        if len(p) < 3:
            params[i, 2] = 1.0
    return N, kinds, params


def get_cond_evap(pdists, s, ξ, ρ_l=1000.0, ctx=None):
    """get_cond_evap(pdists, s, ξ, ρ_l) — src/Sources/Condensation.jl:22-37."""
    ctx = ctx or default_context()
    N, kinds, params = _pack(pdists)
    out = np.zeros(sum(nparams(d) for d in pdists))
    L.check(L.load().cloudy_get_cond_evap_1(ctx.handle, N, kinds, L.dptr(params), float(s), float(ξ), float(ρ_l), L.dptr(out)))
    return tuple(out.tolist())


def get_standard_N_q(pdists, size_cutoff=1e-6, ctx=None):
    """get_standard_N_q(pdists, size_cutoff) → (; N_liq, N_rai, M_liq, M_rai) — ParticleDistributions.jl:634-687."""
    from types import SimpleNamespace
    ctx = ctx or default_context()
    N, kinds, params = _pack(pdists)
    out = np.zeros(4)
    L.check(L.load().cloudy_get_standard_N_q_1(ctx.handle, N, kinds, L.dptr(params), float(size_cutoff), L.dptr(out)))
    return SimpleNamespace(N_liq=out[0], N_rai=out[1], M_liq=out[2], M_rai=out[3])
