"""Batched device state and the ODE right-hand sides that consume the hot path.

``ParcelEnsemble`` holds the prognostic moments of many independent parcels (box model) or of the cells of
many columns (rainshaft) on the GPU, structure-of-arrays.  ``make_box_model_rhs`` / ``make_rainshaft_rhs``
mirror test/examples/utils/box_model_helpers.jl:22-53 and rainshaft_helpers.jl:45-88 for one parcel /
column (the reference's calling convention) and for ensembles."""
import ctypes as C
from types import SimpleNamespace

import numpy as np

from . import _lib as L
from .coalescence import (AnalyticalCoalStyle, CoalescenceData, FixedThreshold, MovingThreshold, apply_config,
                          build_config, new_uid)
from .context import Context, default_context
from .distributions import nparams


class ModelParameters(SimpleNamespace):
    """The drivers' ``ODE_parameters`` named tuple: pdists, coal_data, NProgMoms, norms, dt[, vel, dz]
    (box_gamma_mixture.jl:30-36, rainshaft_gamma_mixture.jl:39-47)."""


class ParcelEnsemble:
    """Device-resident moments of ``n`` parcels / cells; slot order = the reference's flat moment vector."""

    def __init__(self, ctx: Context, n: int):
        self.ctx = ctx
        self.n = int(n)
        h = C.c_void_p()
        L.check(L.load().cloudy_state_create(ctx.handle, self.n, C.byref(h)))
        self.handle = h
        ns = C.c_int32()
        st = C.c_int64()
        p = C.c_void_p()
        L.check(L.load().cloudy_state_device_ptr(self.handle, C.byref(p), C.byref(st), C.byref(ns)))
        self.n_slots, self.stride = ns.value, st.value

    def device_ptr(self):
        p = C.c_void_p()
        L.check(L.load().cloudy_state_device_ptr(self.handle, C.byref(p), None, None))
        return p.value

    def regime_sort(self):
        """Move the parcels into regime order now (cloudy_state_regime_sort); the state is logically unchanged."""
        L.check(L.load().cloudy_state_regime_sort(self.ctx.handle, self.handle))
        return self

    def order(self):
        """Original parcel index of every position (host copy), or None while position == parcel index."""
        p = C.c_void_p()
        L.check(L.load().cloudy_state_order(self.handle, C.byref(p)))
        if not p.value:
            return None
        out = np.empty(self.n, dtype=np.int32)
        self.ctx.sync()
        _DeviceBuffer.memcpy_d2h(out, p.value, 4 * self.n)
        return out

    def upload(self, host):
        a = np.ascontiguousarray(host, dtype=np.float64)
        if a.shape != (self.n, self.n_slots):
            raise ValueError(f"expected host array of shape {(self.n, self.n_slots)}, got {a.shape}")
        L.check(L.load().cloudy_state_upload(self.ctx.handle, self.handle, L.dptr(a), self.n))
        self.ctx.sync()  # the staging buffer is reused; keep `a` alive until the copy has landed
        return self

    def download(self, out=None):
        a = out if out is not None else np.empty((self.n, self.n_slots), dtype=np.float64)
        L.check(L.load().cloudy_state_download(self.ctx.handle, self.handle, L.dptr(a), self.n))
        return a

    def close(self):
        if getattr(self, "handle", None):
            L.load().cloudy_state_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DeviceBuffer:
    """Plain device scratch (n doubles) borrowed from a one-slot... implemented with cudart through ctypes."""

    _rt = None

    @staticmethod
    def _runtime():
        if _DeviceBuffer._rt is None:
            import ctypes.util
            import glob
            cands = glob.glob("/usr/local/cuda/lib64/libcudart.so*") + [ctypes.util.find_library("cudart") or "libcudart.so"]
            _DeviceBuffer._rt = C.CDLL(cands[0])
        return _DeviceBuffer._rt

    def __init__(self, ctx, n):
        _DeviceBuffer._runtime()
        self.ctx = ctx
        self.n = int(n)
        p = C.c_void_p()
        rc = _DeviceBuffer._rt.cudaMalloc(C.byref(p), C.c_size_t(8 * max(self.n, 1)))
        if rc != 0:
            raise L.CloudyError(-2, f"cudaMalloc failed ({rc})")
        self.ptr = p.value

    @staticmethod
    def memcpy_d2h(out, dev_ptr, nbytes):
        _DeviceBuffer._runtime()
        rc = _DeviceBuffer._rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(dev_ptr), C.c_size_t(nbytes), C.c_int(2))
        if rc != 0:
            raise L.CloudyError(-2, f"cudaMemcpy failed ({rc})")

    def download(self, out):
        self.ctx.sync()
        rc = _DeviceBuffer._rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), C.c_size_t(8 * self.n), C.c_int(2))
        if rc != 0:
            raise L.CloudyError(-2, f"cudaMemcpy failed ({rc})")

    def __del__(self):
        try:
            _DeviceBuffer._rt.cudaFree(C.c_void_p(self.ptr))
        except Exception:
            pass


class CoalescenceModel:
    """Run-constant configuration on a device context + the batched operators."""

    def __init__(self, par, ctx: Context = None, nz: int = 1):
        self.ctx = ctx or default_context()
        self.par = par
        self.kinds = tuple(d.kind for d in par.pdists)
        self.NProgMoms = tuple(par.NProgMoms)
        if tuple(nparams(d) for d in par.pdists) != self.NProgMoms:
            raise ValueError("NProgMoms does not match the distributions")
        self.nz = int(nz)
        self.cfg = build_config(self.kinds, par.coal_data, norms=par.norms, vel=getattr(par, "vel", ()),
                                dz=getattr(par, "dz", 1.0), nz=self.nz)
        self._key = ("model", new_uid())
        self.n_slots = sum(self.NProgMoms)
        self.activate()

    def activate(self):
        apply_config(self.ctx, self.cfg, key=self._key)

    def ensemble(self, n: int) -> ParcelEnsemble:
        self.activate()
        return ParcelEnsemble(self.ctx, n)

    # ---- batched operators -----------------------------------------------------------------------
    def coal_tendency(self, m: ParcelEnsemble, dm: ParcelEnsemble):
        self.activate()
        L.check(L.load().cloudy_coal_tendency(self.ctx.handle, m.handle, dm.handle))

    def sedimentation_flux(self, m: ParcelEnsemble, flux: ParcelEnsemble):
        self.activate()
        L.check(L.load().cloudy_sedimentation_flux(self.ctx.handle, m.handle, flux.handle))

    def rainshaft_rhs(self, m: ParcelEnsemble, dm: ParcelEnsemble):
        self.activate()
        L.check(L.load().cloudy_rainshaft_rhs(self.ctx.handle, m.handle, dm.handle))

    def ssprk33_steps(self, u: ParcelEnsemble, dt: float, n_steps: int, model: int = L.MODEL_BOX):
        self.activate()
        L.check(L.load().cloudy_ssprk33_steps(self.ctx.handle, u.handle, float(dt), int(n_steps), int(model)))

    def cond_evap(self, m: ParcelEnsemble, dm: ParcelEnsemble, s: float, ξ: float, ρ_l: float = 1000.0, d_s_ptr: int = 0):
        """rhs_condensation! for every parcel (box_model_helpers.jl:55-67); ``d_s_ptr``: optional device array of per-parcel s."""
        self.activate()
        L.check(L.load().cloudy_cond_evap(self.ctx.handle, m.handle, float(s), C.c_void_p(d_s_ptr) if d_s_ptr else None, float(ξ),
                                          float(ρ_l), dm.handle))

    def standard_N_q(self, m: ParcelEnsemble, size_cutoff: float, normalized: bool = False):
        """get_standard_N_q for every parcel → (4, n) array (N_liq, N_rai, M_liq, M_rai), cf. netcdf_helpers.jl:106-121."""
        self.activate()
        out = np.zeros((4, m.n))
        buf = _DeviceBuffer(self.ctx, 4 * m.n)
        L.check(L.load().cloudy_standard_N_q(self.ctx.handle, m.handle, float(size_cutoff), int(bool(normalized)), C.c_void_p(buf.ptr)))
        buf.download(out)
        return out

    def moment_sums(self, u: ParcelEnsemble):
        out = np.zeros(self.n_slots)
        L.check(L.load().cloudy_moment_sums(self.ctx.handle, u.handle, L.dptr(out)))
        return out

    def moment_sums_allreduce(self, u: ParcelEnsemble, wait: bool = True):
        """Per-slot sums over the parcels of ALL ranks (cloudy_moment_sums_allreduce: local reduction + NCCL all-reduce on a
        side stream).  ``wait=False`` only enqueues; collect the result later with ``moment_sums_fetch``."""
        if wait:
            out = np.zeros(self.n_slots)
            L.check(L.load().cloudy_moment_sums_allreduce(self.ctx.handle, u.handle, L.dptr(out)))
            return out
        L.check(L.load().cloudy_moment_sums_allreduce(self.ctx.handle, u.handle, None))
        return None

    def moment_sums_fetch(self):
        out = np.zeros(self.n_slots)
        L.check(L.load().cloudy_moment_sums_fetch(self.ctx.handle, L.dptr(out)))
        return out

    def moment_sums_device(self, u: ParcelEnsemble, d_out_ptr: int):
        L.check(L.load().cloudy_moment_sums_device(self.ctx.handle, u.handle, C.c_void_p(d_out_ptr)))

    def coal_tendency_host(self, host_m, host_dm=None):
        a = np.ascontiguousarray(host_m, dtype=np.float64)
        out = host_dm if host_dm is not None else np.empty_like(a)
        self.activate()
        L.check(L.load().cloudy_coal_tendency_host(self.ctx.handle, L.dptr(a), L.dptr(out), a.shape[0]))
        return out


def make_box_model_rhs(coal_type, threshold_style=None):
    """make_box_model_rhs(coal_type, threshold_style) — box_model_helpers.jl:22-27.  Returns ``rhs!(dm, m, par, t)``;
    ``m``/``dm`` are one moment vector (the reference's call) or an (n_parcels, n_moments) array."""
    if not isinstance(coal_type, AnalyticalCoalStyle):
        raise ValueError("Invalid coal style!")
    threshold_style = threshold_style or FixedThreshold()

    def rhs(dm, m, par, t):
        if isinstance(threshold_style, MovingThreshold) != isinstance(par.coal_data.threshold_style, MovingThreshold):
            raise ValueError("threshold style does not match coal_data")
        model = getattr(par, "_cloudy_model", None)  # cached on the parameter object itself (id() values are recycled)
        if model is None:
            model = CoalescenceModel(par)
            par._cloudy_model = model
        a = np.atleast_2d(np.asarray(m, dtype=np.float64))
        out = model.coal_tendency_host(a)
        np.asarray(dm)[...] = out.reshape(np.shape(dm))
        return dm

    return rhs


def make_rainshaft_rhs(coal_type):
    """make_rainshaft_rhs(coal_type) — rainshaft_helpers.jl:45-88.  Returns ``rhs(m, p, t)`` for an (nz, nmom)
    column (or (ncol, nz, nmom) columns); ``m`` is clipped at zero IN PLACE like the reference (:52)."""
    if not isinstance(coal_type, AnalyticalCoalStyle):
        raise ValueError("Invalid coal style!")

    def rhs(m, p, t):
        m = np.asarray(m)
        nz = m.shape[-2]
        models = getattr(p, "_cloudy_rain_models", None)
        if models is None:
            models = p._cloudy_rain_models = {}
        model = models.get(nz)
        if model is None:
            model = models[nz] = CoalescenceModel(p, nz=nz)
        flat = np.ascontiguousarray(m.reshape(-1, m.shape[-1]), dtype=np.float64)
        u = model.ensemble(flat.shape[0]).upload(flat)
        du = model.ensemble(flat.shape[0])
        model.rainshaft_rhs(u, du)
        m[...] = u.download().reshape(m.shape)
        return du.download().reshape(m.shape)

    return rhs
