import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
for gen, n in ((W.c2_gamma_exp, 20000), (W.c4_three_modes, 6000)):
    par, state = gen(n_parcels=n)
    model = cb.CoalescenceModel(par)
    u = model.ensemble(n).upload(state); du = model.ensemble(n)
    model.coal_tendency(u, du); ref = du.download()
    model.ctx.set_regime_sort(True)
    model.coal_tendency(u, du); got = du.download()
    model.ctx.set_regime_sort(False)
    diff = got != ref
    rows = np.where(diff.any(axis=1))[0]
    print(gen.__name__, "rows differing", len(rows), "of", n)
    for r in rows[:5]:
        print(r, state[r], "\n  ", ref[r], "\n  ", got[r], "\n  rel", np.abs(got[r]-ref[r])/np.abs(ref[r]).max())
