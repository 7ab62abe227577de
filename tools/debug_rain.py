import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
from oracle import cloudy_oracle as O
from tests.oracle_bridge import oracle_params
from tests.test_gpu_parity import _rain_state
nz, ncol = 20, 3
par, cols = W.c3_rainshaft(n_columns=ncol, nz=nz)
st = _rain_state(cols)
opar = oracle_params(par)
flat = st.reshape(-1, 6).copy()
flat[flat < 0] = 0
model = cb.CoalescenceModel(par, nz=nz)
u = model.ensemble(flat.shape[0]).upload(flat)
d = model.ensemble(flat.shape[0])
model.coal_tendency(u, d); coal = d.download()
model.sedimentation_flux(u, d); flux = d.download()
model.rainshaft_rhs(u, d); rhs = d.download()
np.set_printoptions(linewidth=200, precision=6)
for i in range(flat.shape[0]):
    pd, mn, norms = O.dists_from_state(flat[i], opar)
    empty = all(v < O.EPS for v in mn)
    rc = np.zeros(6) if empty else O.rhs_coal(flat[i], opar)
    rf = O.sedimentation_flux_state(flat[i], opar)
    ec = np.max(np.abs(coal[i] - rc) / (np.abs(rc) + 1e-300)) if not empty else 0
    ef = np.max(np.abs(flux[i] - rf) / (np.abs(rf) + 1e-300))
    if ec > 1e-9 or ef > 1e-9:
        print(i, "coal err", ec, "flux err", ef, "state", flat[i], "\n   got", coal[i], "\n   ref", rc, "\n   dists", pd)
full = O.rainshaft_rhs(st[0].copy(), opar)
print("rhs col0 maxerr rows:", np.max(np.abs(rhs[:nz] - full) / (np.abs(full) + 1e-300), axis=1))
