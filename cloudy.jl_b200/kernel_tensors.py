"""Kernels — host mirror of src/Kernels/KernelFunctions.jl and src/Kernels/KernelTensors.jl.

Init-time only (the reference says so itself, KernelTensors.jl:77); nothing here is on the GPU path.
The tensor VALUES are the hot-path input."""
import math

import numpy as np

EPS = float(np.finfo(np.float64).eps)


# ---- KernelFunctions.jl:39-116 ----------------------------------------------------------------------
class CoalescenceKernelFunction:
    pass


class ConstantKernelFunction(CoalescenceKernelFunction):
    def __init__(self, coll_coal_rate):
        self.coll_coal_rate = float(coll_coal_rate)

    def __call__(self, x, y):
        return self.coll_coal_rate

    def normalized(self, norms):  # :124-129
        return ConstantKernelFunction(self.coll_coal_rate * norms[0])


class LinearKernelFunction(CoalescenceKernelFunction):
    def __init__(self, coll_coal_rate):
        self.coll_coal_rate = float(coll_coal_rate)

    def __call__(self, x, y):
        return self.coll_coal_rate * (x + y)

    def normalized(self, norms):  # :131-136
        return LinearKernelFunction(self.coll_coal_rate * norms[0] * norms[1])


class HydrodynamicKernelFunction(CoalescenceKernelFunction):
    def __init__(self, coal_eff):
        self.coal_eff = float(coal_eff)

    def __call__(self, x, y):  # :102-108
        r1 = (3 / 4 / math.pi * x) ** (1 / 3)
        r2 = (3 / 4 / math.pi * y) ** (1 / 3)
        A1 = math.pi * r1 ** 2
        A2 = math.pi * r2 ** 2
        return self.coal_eff * (r1 + r2) ** 2 * abs(A1 - A2)

    def normalized(self, norms):  # :138-143
        return HydrodynamicKernelFunction(self.coal_eff * norms[0] * norms[1] ** (4 / 3))


class LongKernelFunction(CoalescenceKernelFunction):
    def __init__(self, x_threshold, coal_rate_below_threshold, coal_rate_above_threshold):
        self.x_threshold = float(x_threshold)
        self.coal_rate_below_threshold = float(coal_rate_below_threshold)
        self.coal_rate_above_threshold = float(coal_rate_above_threshold)

    def __call__(self, x, y):  # :110-116
        if x < self.x_threshold and y < self.x_threshold:
            return self.coal_rate_below_threshold * (x ** 2 + y ** 2)
        return self.coal_rate_above_threshold * (x + y)

    def normalized(self, norms):  # :145-154
        return LongKernelFunction(self.x_threshold / norms[1], self.coal_rate_below_threshold * norms[0] * norms[1] ** 2,
                                  self.coal_rate_above_threshold * norms[0] * norms[1])


def get_normalized_kernel_func(kern, norms):
    return kern.normalized(norms)


# ---- KernelTensors.jl ---------------------------------------------------------------------------------
def check_symmetry(obj, n_test=1000):
    """KernelTensors.jl:157-181 (array or function)."""
    if callable(obj):
        rng = np.random.default_rng(0)
        t = rng.random((n_test, 2))
        for a, b in t:
            if abs(obj(a, b) - obj(b, a)) > 1e-6:
                raise ValueError("function likely not symmetric.")
        return
    arr = np.asarray(obj)
    if arr.size > 1:
        if arr.ndim != 2 or arr.shape[0] != arr.shape[1]:
            raise ValueError("array needs to be quadratic in order to be symmetric.")
        n = arr.shape[0]
        for i in range(n):
            for j in range(i + 1, n):
                if arr[i, j] != arr[j, i]:
                    raise ValueError("array not symmetric.")


def polyfit(kernel_func, r: int, limit: float, lower_limit: float = 0.0, norms=(1e6, 1e-9), npoints: int = 10):
    """KernelTensors.jl:78-146.  Same sample grid, same pinned C11 and the same symmetric monomial basis; the
    2-norm loss is minimised exactly by linear least squares instead of the reference's Nelder-Mead iteration
    (Optim.jl, un-vendored), so the result is the minimiser that iteration approximates."""
    if isinstance(kernel_func, CoalescenceKernelFunction):
        kf = get_normalized_kernel_func(kernel_func, norms)
    else:
        kf = kernel_func
        norms = (1.0, 1.0)
    limit_n = limit / norms[1]
    lower_n = lower_limit / norms[1]
    check_symmetry(kf)
    if limit_n <= lower_n or lower_n < 0:
        raise ValueError("polyfit limits improperly specified")
    d = limit_n / (npoints - 1)
    idx = np.arange(npoints * npoints)
    x_ = (idx % npoints) * d
    y_ = np.floor(idx / npoints) * d
    keep = (y_ >= lower_n) & (y_ - x_ >= 0)
    x, y = x_[keep], y_[keep]
    C11 = max(EPS, kf(0.0, 0.0))
    if r == 0:
        return np.array([[C11 / norms[0]]])
    # the reference's loss uses the FULL tensor grid z[i][j] = K(x_i, y_j) over the kept x's and y's (:131-135)
    X, Y = np.meshgrid(x, y, indexing="ij")
    Z = np.vectorize(kf)(X, Y) - C11
    cols, pairs = [], []
    for j in range(r + 1):
        for i in range(j + 1):
            if i == 0 and j == 0:
                continue
            basis = X ** i * Y ** j if i == j else (X ** i * Y ** j + X ** j * Y ** i)
            cols.append(basis.ravel())
            pairs.append((i, j))
    A = np.stack(cols, axis=1)
    scale = np.linalg.norm(A, axis=0)
    scale[scale == 0] = 1.0
    sol = np.linalg.lstsq(A / scale, Z.ravel(), rcond=None)[0] / scale
    Cm = np.zeros((r + 1, r + 1))
    Cm[0, 0] = C11
    for (i, j), v in zip(pairs, sol):
        Cm[i, j] = v
        Cm[j, i] = v
    out = np.empty_like(Cm)
    for i in range(r + 1):
        for j in range(r + 1):
            out[i, j] = Cm[i, j] / (norms[0] * norms[1] ** float(i + j))
    return out


class CoalescenceTensor:
    """CoalescenceTensor{P,FT} — KernelTensors.jl:44-64.  ``CoalescenceTensor(c)`` from a symmetric P×P array, or
    ``CoalescenceTensor(kernel_func, order, limit[, lower_limit, norms])`` by polynomial fit."""

    def __init__(self, c_or_func, order=None, limit=None, lower_limit=0.0, norms=(1e6, 1e-9)):
        if order is None:
            c = np.array(c_or_func, dtype=np.float64)
            if c.ndim == 1 and c.size == 1:
                c = c.reshape(1, 1)
            check_symmetry(c)
            self.c = c
        else:
            self.c = polyfit(c_or_func, int(order), float(limit), float(lower_limit), norms)

    @property
    def P(self):
        return self.c.shape[0]


def get_normalized_kernel_tensor(kernel: CoalescenceTensor, norms) -> CoalescenceTensor:
    """KernelTensors.jl:189-199"""
    P = kernel.P
    c = np.empty((P, P))
    for i in range(P):
        for j in range(P):
            c[i, j] = kernel.c[i, j] * (norms[0] * norms[1] ** float(i + j))
    return CoalescenceTensor(c)
