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
