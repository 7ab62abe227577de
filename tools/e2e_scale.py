"""e2e leg: time of cloudy_coal_tendency_host vs ensemble size (slope = steady state, intercept = fill/drain + call overhead)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
nmax = 1 << 23
par, state = W.c2_gamma_exp(nmax)
model = cb.CoalescenceModel(par)
hin = torch.from_numpy(state).pin_memory().numpy(); hout = torch.empty((nmax, 5), dtype=torch.float64).pin_memory().numpy()
for n in (1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    for _ in range(3): model.coal_tendency_host(hin[:n], hout[:n])
    t0 = time.perf_counter()
    for _ in range(10): model.coal_tendency_host(hin[:n], hout[:n])
    dt = (time.perf_counter() - t0) / 10
    print(n, f"{dt*1e3:.3f} ms  {n/dt:.3e} parcel/s  {2*40*n/dt/1e9:.1f} GB/s both ways")
