import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
np.set_printoptions(linewidth=200)
par, state = W.c2_gamma_exp(n_parcels=20000)
model = cb.CoalescenceModel(par)
n = state.shape[0]
for steps in (1, 2):
    u = model.ensemble(n).upload(state)
    model.ssprk33_steps(u, 0.05, steps, cb.MODEL_BOX)
    out = u.download()
    bad = np.where(~np.isfinite(out).all(axis=1))[0]
    print("steps", steps, "bad rows", len(bad))
    for r in bad[:4]:
        print(r, state[r], out[r])
du = model.ensemble(n); u = model.ensemble(n).upload(state)
model.coal_tendency(u, du); t = du.download()
print("tendency nonfinite rows", (~np.isfinite(t).all(axis=1)).sum())
ratio = np.abs(t) * 0.05 / (np.abs(state) + 1e-300)
print("max dt*|f|/|u| per slot", ratio.max(axis=0))
