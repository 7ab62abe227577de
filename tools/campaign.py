"""One-off accuracy campaign: thread-per-parcel kernel vs lane-cooperative kernel (independent CUDA implementations:
the latter has no Taylor zone, two exponentials, library exp) on a large wide-parameter ensemble; worst offenders are
re-checked against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
rng = np.random.default_rng(99)
if which == "c2":
    par, _ = W.c2_gamma_exp(n_parcels=8)
    k = np.exp(rng.uniform(np.log(1e-3), np.log(12.0), n)); th = np.exp(rng.uniform(np.log(1e-3), np.log(1e2), n))
    nn = np.exp(rng.uniform(np.log(1e-3), np.log(1e3), n))
    m1 = np.stack([nn, nn * k * th, nn * k * (k + 1) * th ** 2], axis=1)
    n2 = np.exp(rng.uniform(np.log(1e-6), 0.0, n)); th2 = np.exp(rng.uniform(0.0, np.log(30.0), n))
    state = np.concatenate([m1, np.stack([n2, n2 * th2], axis=1)], axis=1) * np.array([1e6, 1e-3, 1e-12, 1e6, 1e-3])
else:
    par, state = W.c4_three_modes(n)
model = cb.CoalescenceModel(par)
model.ctx.set_lanes(1); a = model.coal_tendency_host(state)
model.ctx.set_lanes(8); b = model.coal_tendency_host(state)
rowmax = np.abs(b).max(axis=1, keepdims=True)
ratio = np.abs(a - b) / (np.abs(b) + 1e-9 * rowmax + 1e-300)
worst = np.argsort(ratio.max(axis=1))[::-1][:40]
print("max ratio", ratio.max(), "frac rows > 1e-9:", (ratio.max(axis=1) > 1e-9).mean())
opar = oracle_params(par)
nbad = 0
for i in worst:
    ref, sc = O.rhs_coal(state[i], opar, return_scale=True)
    tol = np.maximum(np.abs(ref), sc)
    ea = np.max(np.abs(a[i] - ref) / tol); eb = np.max(np.abs(b[i] - ref) / tol)
    flag = "BAD" if (ea > 1e-9 or eb > 1e-9) else "ok"
    nbad += flag == "BAD"
    if flag == "BAD" or i in worst[:5]:
        print(f"i={i} ratio={ratio[i].max():.2e} tpp_err={ea:.2e} lanes8_err={eb:.2e} {flag} state={state[i]}")
print("bad among worst 40:", nbad)
