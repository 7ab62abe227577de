"""Flat moment-vector layout helpers — mirror of src/helper_functions.jl:13-58 (1-based like Julia)."""
from typing import Sequence, Tuple


def get_dist_moment_ind(NProgMoms: Sequence[int], i: int, m: int) -> int:
    """helper_functions.jl:13-21."""
    if i < 1 or i > len(NProgMoms):
        raise IndexError("distribution index out of range")
    if not (0 < m <= NProgMoms[i - 1]):
        raise ValueError(
            "moment index must be positive integer and equal or smaller than the dist number of prognostic moments!!!"
        )
    return m if i == 1 else sum(NProgMoms[: i - 1]) + m


def get_dist_moments_ind_range(NProgMoms: Sequence[int], i: int) -> range:
    """helper_functions.jl:29-33 (inclusive 1-based range)."""
    if i < 1 or i > len(NProgMoms):
        raise IndexError("distribution index out of range")
    last_ind = 0 if i == 1 else sum(NProgMoms[: i - 1])
    return range(last_ind + 1, last_ind + NProgMoms[i - 1] + 1)


def get_moments_normalizing_factors(NProgMoms: Sequence[int], norms: Tuple[float, float]):
    """helper_functions.jl:40-53."""
    if norms[0] <= 0 or norms[1] <= 0:
        raise ValueError("norms must be positive!")
    return tuple(norms[0] * norms[1] ** (j - 1) for n_i in NProgMoms for j in range(1, n_i + 1))


def rflatten(tup):
    """helper_functions.jl:55-58."""
    out = []
    for x in tup:
        if isinstance(x, (tuple, list)):
            out.extend(rflatten(x))
        else:
            out.append(x)
    return tuple(out)
