"""Accuracy campaign for the MovingThreshold path: wide-parameter Gamma + Exponential ensemble against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
worst_all = 0.0
for pct in (0.97, 0.5, 0.999, 0.05):
    rng = np.random.default_rng(int(pct * 1000))
    par, _ = W.moving_gamma_exp(n_parcels=8, percentile=pct)
    k = np.exp(rng.uniform(np.log(1e-2), np.log(10.0), n)); th = np.exp(rng.uniform(np.log(1e-3), np.log(1e3), n))
    nn = np.exp(rng.uniform(np.log(1e-3), np.log(1e3), n))
    m1 = np.stack([nn, nn * k * th, nn * k * (k + 1) * th ** 2], axis=1)
    n2 = np.exp(rng.uniform(np.log(1e-6), 0.0, n)); th2 = np.exp(rng.uniform(0.0, np.log(30.0), n))
    state = np.concatenate([m1, np.stack([n2, n2 * th2], axis=1)], axis=1) * np.array([1e6, 1e-3, 1e-12, 1e6, 1e-3])
    model = cb.CoalescenceModel(par)
    got = model.coal_tendency_host(state)
    opar = oracle_params(par)
    errs = np.zeros(n)
    for i in range(n):
        ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
        errs[i] = np.max(np.abs(got[i] - ref) / np.maximum(np.maximum(np.abs(ref), sc), 1e-300))
    w = int(np.argmax(errs))
    print(f"percentile {pct}: worst {errs[w]:.3e} (k={k[w]:.4g}, theta={th[w]:.4g}), rows > 1e-9: {(errs > 1e-9).sum()} of {n}")
    worst_all = max(worst_all, errs[w])
print("worst overall", worst_all)
