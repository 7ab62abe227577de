import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
par, cols = W.c3_rainshaft(n_columns=1, nz=20)
model = cb.CoalescenceModel(par, nz=20)
u = model.ensemble(20).upload(cols[0])
print("direct:", model.standard_N_q(u, 5.236e-10, normalized=False)[:, 8:13])
model.ssprk33_steps(u, par.dt, 20, cb.MODEL_RAINSHAFT)
st = u.download()
print("state rows", st[8:13])
print("after steps:", model.standard_N_q(u, 5.236e-10, normalized=False)[:, 8:13])
u3 = model.ensemble(60).upload(np.concatenate([cols[0], st, st]))
print("n=60:", model.standard_N_q(u3, 5.236e-10, normalized=False)[:, 28:33])
