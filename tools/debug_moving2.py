import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params
from tests.test_gpu_parity import _moving_mixed
from scipy.special import gammaincinv
for pct, klo, khi, wi in ((0.97, 0.03, 0.4, 29), (0.01, 0.5, 5.0, 2)):
    par, state = _moving_mixed(cb, 256, 103, klo=klo, khi=khi, percentile=pct)
    mn = state[wi] / np.array([1e6, 1e-3, 1e-12, 1e6, 1e-3])
    mean = mn[1] / mn[0]; k = mean / (mn[2] / mn[1] - mean); th = mean / k
    X = gammaincinv(k, pct)
    dev_thr = cb.compute_threshold(cb.GammaPrimitiveParticleDistribution(mn[0], th, k), pct)
    print("pct", pct, "k", k, "thr scipy", th * X, "device ref", dev_thr, "rel", dev_thr / (th * X) - 1)
    # moving on device vs oracle for this parcel alone, and fixed threshold at scipy's value
    opar = oracle_params(par)
    model = cb.CoalescenceModel(par)
    got = model.coal_tendency_host(state[wi:wi + 1])[0]
    ref, sc = O.rhs_coal(state[wi], opar, return_scale=True)
    print("  moving: dev", got, "\n          ref", ref, "\n   err", np.abs(got - ref) / np.maximum(np.abs(ref), sc))
    cdf = cb.CoalescenceData(W.linear_tensor(5.0), par.NProgMoms, (th * X * 1e-9, math.inf), W.NORMS)
    parf = type(par)(**{**vars(par), "coal_data": cdf})
    gotf = cb.CoalescenceModel(parf).coal_tendency_host(state[wi:wi + 1])[0]
    reff, scf = O.rhs_coal(state[wi], oracle_params(parf), return_scale=True)
    print("  fixed:  err", np.abs(gotf - reff) / np.maximum(np.abs(reff), scf), " fixed-ref vs moving-ref", np.abs(reff - ref) / np.maximum(np.abs(ref), sc))
