"""ctypes binding of libcloudy_b200.so (the C ABI declared in include/cloudy_b200.h).

This is the in-repo stand-in for the Julia ``ccall`` wrapper (julia/CloudyB200.jl): the same symbols,
the same argument order.  There is no CPU fallback — if the library is missing, or no CUDA device is
present when a compute call is made, an exception is raised."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLOUDY_LIB") or os.path.join(HERE, "libcloudy_b200.so")  # CLOUDY_LIB: development variants (build.py)

MAX_MODES, MAX_P, MAX_VEL, MAX_SLOTS, MAX_NODES = 4, 5, 4, 12, 512
EXPONENTIAL, GAMMA, LOGNORMAL, MONODISPERSE = 0, 1, 2, 3
FIXED_THRESHOLD, MOVING_THRESHOLD = 0, 1
MODEL_BOX, MODEL_RAINSHAFT = 0, 1


class CloudyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libcloudy_b200 error {code}: {msg}")
        self.code = code


class cloudy_config(C.Structure):
    _fields_ = [
        ("n_modes", C.c_int32), ("P", C.c_int32),
        ("kind", C.c_int32 * MAX_MODES), ("nprog", C.c_int32 * MAX_MODES),
        ("threshold_style", C.c_int32), ("n_mom_max", C.c_int32),
        ("n_2d_ints", C.c_int32 * MAX_MODES), ("n_bins", C.c_int32 * MAX_MODES),
        ("n_vel", C.c_int32), ("nz", C.c_int32), ("bins_per_log_unit", C.c_int32), ("reserved", C.c_int32),
        ("c", (((C.c_double * MAX_P) * MAX_P) * MAX_MODES) * MAX_MODES),
        ("thresholds", C.c_double * MAX_MODES), ("x_min", C.c_double * MAX_MODES), ("dx", C.c_double * MAX_MODES),
        ("norms", C.c_double * 2), ("k_range", C.c_double * 2),
        ("vel", (C.c_double * 2) * MAX_VEL), ("dz", C.c_double),
    ]


_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I32 = C.POINTER(C.c_int32)
_I64 = C.POINTER(C.c_int64)

# every symbol include/cloudy_b200.h declares: name → (argtypes)
SIGNATURES = {
    "cloudy_ctx_create": (C.c_int, _P, C.POINTER(_P)),
    "cloudy_ctx_destroy": (_P,),
    "cloudy_config_set": (_P, C.POINTER(cloudy_config)),
    "cloudy_sync": (_P,),
    "cloudy_launch_count": (_P, _I64),
    "cloudy_set_lanes": (_P, C.c_int),
    "cloudy_set_regime_sort": (_P, C.c_int),
    "cloudy_state_create": (_P, C.c_int64, C.POINTER(_P)),
    "cloudy_state_destroy": (_P,),
    "cloudy_state_upload": (_P, _P, _D, C.c_int64),
    "cloudy_state_download": (_P, _P, _D, C.c_int64),
    "cloudy_state_device_ptr": (_P, C.POINTER(C.c_void_p), _I64, _I32),
    "cloudy_state_copy": (_P, _P, _P),
    "cloudy_state_regime_sort": (_P, _P),
    "cloudy_state_order": (_P, C.POINTER(C.c_void_p)),
    "cloudy_set_resort_interval": (_P, C.c_int32),
    "cloudy_sort_count": (_P, _I64),
    "cloudy_coal_tendency": (_P, _P, _P),
    "cloudy_sedimentation_flux": (_P, _P, _P),
    "cloudy_rainshaft_rhs": (_P, _P, _P),
    "cloudy_ssprk33_steps": (_P, _P, C.c_double, C.c_int32, C.c_int32),
    "cloudy_moment_sums_device": (_P, _P, C.c_void_p),
    "cloudy_moment_sums": (_P, _P, _D),
    "cloudy_comm_unique_id": (C.c_void_p,),
    "cloudy_comm_init": (_P, C.c_int32, C.c_int32, C.c_void_p),
    "cloudy_comm_destroy": (_P,),
    "cloudy_comm_info": (_P, _I32, _I32, _I32),
    "cloudy_moment_sums_allreduce": (_P, _P, _D),
    "cloudy_moment_sums_fetch": (_P, _D),
    "cloudy_cond_evap": (_P, _P, C.c_double, C.c_void_p, C.c_double, C.c_double, _P),
    "cloudy_standard_N_q": (_P, _P, C.c_double, C.c_int32, C.c_void_p),
    "cloudy_coal_tendency_host": (_P, _D, _D, C.c_int64),
    "cloudy_error_count": (_P, _I64),
    "cloudy_moment": (_P, C.c_int32, _D, C.c_double, _D),
    "cloudy_update_dist_from_moments": (_P, C.c_int32, _D, _D, _D, _I32),
    "cloudy_moment_source_helper": (_P, C.c_int32, _D, C.c_double, C.c_double, C.c_double, C.c_int32, _D),
    "cloudy_compute_threshold": (_P, C.c_int32, _D, C.c_double, C.c_double, _D),
    "cloudy_get_coal_ints_1": (_P, _D, _D),
    "cloudy_get_sedimentation_flux_1": (_P, C.c_int32, _I32, _D, C.c_int32, _D, _D),
    "cloudy_get_cond_evap_1": (_P, C.c_int32, _I32, _D, C.c_double, C.c_double, C.c_double, _D),
    "cloudy_get_standard_N_q_1": (_P, C.c_int32, _I32, _D, C.c_double, _D),
    "cloudy_integrate_simpson": (_P, C.c_int32, C.c_double, _D, _D),
    "cloudy_measure_fp64_peak": (_P, _D),
    "cloudy_config_sizeof": (),
    "cloudy_config_offsets": (_I64, C.c_int32),
}

_lib = None


def load():
    """Load the shared library (raises if it has not been built — there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  cloudy_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = list(argtypes)
        fn.restype = C.c_int
    lib.cloudy_last_error.argtypes = []
    lib.cloudy_last_error.restype = C.c_char_p
    lib.cloudy_config_sizeof.restype = C.c_int64
    lib.cloudy_config_offsets.restype = C.c_int32
    check_config_layout(lib)
    _lib = lib
    return lib


def config_layout():
    """(sizeof, [field offsets]) of this module's mirror of ``cloudy_config``"""
    return C.sizeof(cloudy_config), [getattr(cloudy_config, name).offset for name, _ in cloudy_config._fields_]


def check_config_layout(lib):
    """the library reports sizeof(cloudy_config) and its field offsets as compiled (cloudy_config_sizeof / cloudy_config_offsets):
    refuse to run with a mirror that has drifted"""
    size, offs = config_layout()
    got = (C.c_int64 * 64)()
    n = lib.cloudy_config_offsets(got, 64)
    if lib.cloudy_config_sizeof() != size or n != len(offs) or list(got[:n]) != offs:
        raise ImportError(f"cloudy_config layout mismatch: library sizeof {lib.cloudy_config_sizeof()} offsets {list(got[:n])}, "
                          f"Python mirror sizeof {size} offsets {offs}")


def check(rc):
    if rc != 0:
        raise CloudyError(rc, load().cloudy_last_error().decode("utf-8", "replace"))


def dptr(arr):
    return arr.ctypes.data_as(_D)
