"""Device context handling for the host mirror."""
import ctypes as C
import threading

from . import _lib as L

_default = None
_lock = threading.Lock()


class Context:
    """Owns a ``cloudy_ctx`` (one device, one stream)."""

    def __init__(self, device: int = 0, stream=None):
        lib = L.load()
        self._lib = lib
        h = C.c_void_p()
        L.check(lib.cloudy_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.device = device
        self.config = None  # the CoalescenceData / model parameters currently on the device

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cloudy_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        L.check(self._lib.cloudy_sync(self.handle))

    def set_lanes(self, lanes: int):
        L.check(self._lib.cloudy_set_lanes(self.handle, int(lanes)))

    def set_regime_sort(self, mode):
        """0/False off, 1/True always, 2 auto (default)"""
        L.check(self._lib.cloudy_set_regime_sort(self.handle, int(mode)))

    def set_resort_interval(self, steps: int):
        """fused time steps between two refreshes of a resident ensemble's regime order (default 10)"""
        L.check(self._lib.cloudy_set_resort_interval(self.handle, int(steps)))

    def sort_count(self) -> int:
        v = C.c_int64()
        L.check(self._lib.cloudy_sort_count(self.handle, C.byref(v)))
        return v.value

    def launch_count(self) -> int:
        v = C.c_int64()
        L.check(self._lib.cloudy_launch_count(self.handle, C.byref(v)))
        return v.value

    def error_count(self) -> int:
        v = C.c_int64()
        L.check(self._lib.cloudy_error_count(self.handle, C.byref(v)))
        return v.value

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        L.check(self._lib.cloudy_measure_fp64_peak(self.handle, C.byref(v)))
        return v.value


def default_context() -> Context:
    global _default
    with _lock:
        if _default is None:
            _default = Context(0)
        return _default
