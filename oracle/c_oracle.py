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
