import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
n = int(sys.argv[1]); which = sys.argv[2]
par, state = (W.moving_gamma_exp if which == "ge" else W.moving_four_modes)(n)
model = cb.CoalescenceModel(par)
if len(sys.argv) > 3: model.ctx.set_regime_sort(int(sys.argv[3]))
u = model.ensemble(n).upload(state); du = model.ensemble(n)
model.coal_tendency(u, du); model.ctx.sync()
print(which, n, "ok", float(np.abs(du.download()).sum()))
