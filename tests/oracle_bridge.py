"""Translate the host mirror's model parameters into the oracle's (test infrastructure)."""
import numpy as np

from oracle import cloudy_oracle as O


def oracle_params(par) -> O.ModelParams:
    cd = par.coal_data
    kinds = tuple(d.kind for d in par.pdists)
    ocd = O.CoalescenceData(cd.N_mom_max, tuple(cd.N_2d_ints), tuple(cd.dist_thresholds),
                            [[np.array(cd.kernels[j][k].c) for k in range(cd.N)] for j in range(cd.N)], cd.P, cd.N,
                            moving=type(cd.threshold_style).__name__ == "MovingThreshold")
    return O.ModelParams(kinds, ocd, tuple(par.NProgMoms), tuple(par.norms), tuple(getattr(par, "vel", ())), getattr(par, "dz", 1.0))


def tendency_close(got, ref, scale, rtol):
    """|Δ| <= rtol * max(|ref|, Σ|terms|) — tendencies contain exact cancellations."""
    got, ref, scale = np.asarray(got), np.asarray(ref), np.asarray(scale)
    tol = rtol * np.maximum(np.abs(ref), scale)
    err = np.abs(got - ref)
    bad = ~((err <= tol) | ((got != got) & (ref != ref)))
    worst = float(np.max(np.where(tol > 0, err / np.where(tol > 0, tol, 1.0), 0.0))) * rtol if got.size else 0.0
    return not bad.any(), worst
