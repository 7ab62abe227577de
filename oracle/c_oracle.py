"""ctypes driver of oracle/liboracle.so (TEST INFRASTRUCTURE / CPU BASELINE ONLY — see cloudy_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


def _host_id() -> str:
    """identifies the CPU the library was compiled for (-march=native): model name + ISA flags"""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        keep = [l for l in txt.splitlines() if l.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("\n".join(keep).encode()).hexdigest()
    except Exception:
        return "unknown"


def build(force=False):
    src = os.path.join(HERE, "cloudy_oracle.c")
    tag = LIB + ".host"
    same_host = os.path.exists(tag) and open(tag).read().strip() == _host_id()
    if force or not os.path.exists(LIB) or not same_host or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "Makefile"))):
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
        with open(tag, "w") as f:
            f.write(_host_id())
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        build()  # no-op when the library exists and was compiled on this host's CPU
        _lib = C.CDLL(LIB)
        _lib.cloudy_oracle_build_flags.restype = C.c_char_p
        _lib.cloudy_oracle_rhs_coal_batch.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64, C.c_int]
        _lib.cloudy_oracle_rhs_coal_batch.restype = C.c_int
        _lib.cloudy_oracle_max_threads.restype = C.c_int
        D = C.POINTER(C.c_double)
        _lib.cloudy_oracle_sedimentation_flux_batch.argtypes = [C.c_void_p, D, D, C.c_int64]
        _lib.cloudy_oracle_sedimentation_flux_batch.restype = C.c_int
        _lib.cloudy_oracle_rainshaft_rhs.argtypes = [C.c_void_p, D, D, C.c_int64, C.c_int]
        _lib.cloudy_oracle_rainshaft_rhs.restype = C.c_int
        _lib.cloudy_oracle_ssprk33.argtypes = [C.c_void_p, D, C.c_double, C.c_int, C.c_int, C.c_int64, C.c_int]
        _lib.cloudy_oracle_ssprk33.restype = C.c_int
        _lib.cloudy_oracle_moment_source_helper.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double,
                                                            C.c_int, C.c_double, C.c_double]
        _lib.cloudy_oracle_moment_source_helper.restype = C.c_double
    return _lib


def build_flags() -> str:
    return load().cloudy_oracle_build_flags().decode()


def max_threads() -> int:
    return load().cloudy_oracle_max_threads()


def rhs_coal_batch(cfg_struct, states, n_threads=0):
    """cfg_struct: the ctypes ``cloudy_config`` (same layout as include/cloudy_b200.h); states (n, n_slots)."""
    a = np.ascontiguousarray(states, dtype=np.float64)
    out = np.empty_like(a)
    rc = load().cloudy_oracle_rhs_coal_batch(C.byref(cfg_struct), a.ctypes.data_as(C.POINTER(C.c_double)),
                                             out.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0], int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return out


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def sedimentation_flux_batch(cfg_struct, states):
    """get_sedimentation_flux .* mom_norms of every cell (Sedimentation.jl:22-37, rainshaft_helpers.jl:74-77); states (n, n_slots)."""
    a = np.ascontiguousarray(states, dtype=np.float64)
    out = np.empty_like(a)
    rc = load().cloudy_oracle_sedimentation_flux_batch(C.byref(cfg_struct), _dp(a), _dp(out), a.shape[0])
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return out


def rainshaft_rhs(cfg_struct, columns, n_threads=0):
    """rainshaft_helpers.jl:47-88 for columns (n_columns, nz, n_slots); ``columns`` is clipped IN PLACE like the reference."""
    if not (columns.flags.c_contiguous and columns.dtype == np.float64):
        raise ValueError("columns must be a C-contiguous float64 array (it is clipped in place)")
    out = np.empty_like(columns)
    rc = load().cloudy_oracle_rainshaft_rhs(C.byref(cfg_struct), _dp(columns), _dp(out), columns.shape[0], int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return out


def ssprk33(cfg_struct, u0, dt, n_steps, model=0, n_threads=0):
    """n_steps of SSPRK33 on a copy of u0: model 0 = box (u0 (n, n_slots)), model 1 = rainshaft (u0 (n_columns, nz, n_slots))."""
    u = np.array(u0, dtype=np.float64, order="C", copy=True)
    rc = load().cloudy_oracle_ssprk33(C.byref(cfg_struct), _dp(u), float(dt), int(n_steps), int(model), u.shape[0], int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return u
