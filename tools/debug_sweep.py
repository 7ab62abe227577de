import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params
par, _ = W.c2_gamma_exp(n_parcels=8)
rng = np.random.default_rng(2718)
n = 600
k = np.exp(rng.uniform(np.log(1e-3), np.log(10.0), n)); k[:20] = 10.0
th = np.exp(rng.uniform(np.log(1e-4), np.log(1e4), n)); th[20:60] = 0.5 / rng.uniform(17.0, 27.0, 40)
nn = np.exp(rng.uniform(np.log(1e-3), np.log(1e3), n))
m1 = np.stack([nn, nn * k * th, nn * k * (k + 1) * th ** 2], axis=1)
n2 = np.exp(rng.uniform(np.log(1e-6), 0.0, n)); th2 = np.exp(rng.uniform(0.0, np.log(30.0), n))
state = np.concatenate([m1, np.stack([n2, n2 * th2], axis=1)], axis=1) * np.array([1e6, 1e-3, 1e-12, 1e6, 1e-3])
opar = oracle_params(par)
model = cb.CoalescenceModel(par)
for lanes in (1, 8):
    model.ctx.set_lanes(lanes)
    got = model.coal_tendency_host(state)
    bad = []
    for i in range(n):
        ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
        tol = np.maximum(np.abs(ref), sc)
        e = np.max(np.abs(got[i] - ref) / tol)
        if not (e <= 1e-9): bad.append((e, i))
    bad.sort(reverse=True)
    print("lanes", lanes, "n bad", len(bad))
    for e, i in bad[:12]:
        print(f"  i={i} err={e:.2e} k={k[i]:.4g} theta={th[i]:.4g} X={0.5/th[i]:.4g} n={nn[i]:.3g}")
