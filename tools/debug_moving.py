import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params, tendency_close
from tests.test_gpu_parity import _moving_mixed
from scipy.special import gammaincinv
for pct, klo, khi in ((0.97, 0.03, 0.4), (0.999, 0.5, 5.0), (0.2, 0.5, 5.0), (0.01, 0.5, 5.0)):
    par, state = _moving_mixed(cb, 256, 103, klo=klo, khi=khi, percentile=pct)
    opar = oracle_params(par)
    model = cb.CoalescenceModel(par)
    got = model.coal_tendency_host(state)
    worst = 0; wi = -1
    for i in range(40):
        ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
        err = np.max(np.abs(got[i] - ref) / np.maximum(np.abs(ref), sc))
        if err > worst: worst, wi = err, i
    mn = state[wi] / np.array([1e6, 1e-3, 1e-12, 1e6, 1e-3])
    mean = mn[1] / mn[0]; k = mean / (mn[2] / mn[1] - mean); th = mean / k
    X = gammaincinv(k, pct)
    print(pct, "worst", worst, "parcel", wi, "k", k, "theta", th, "X", X, "thr", th * X)
