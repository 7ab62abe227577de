import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
n = 1 << 23
par, state = W.c2_gamma_exp(n)
model = cb.CoalescenceModel(par)
hin = torch.from_numpy(state).pin_memory().numpy(); hout = torch.empty((n, 5), dtype=torch.float64).pin_memory().numpy()
for chunk in (65536, 131072, 262144, 524288, 1048576):
    os.environ["CLOUDY_PIPE_CHUNK"] = str(chunk)
    for _ in range(2): model.coal_tendency_host(hin, hout)
    t0 = time.perf_counter()
    for _ in range(5): model.coal_tendency_host(hin, hout)
    dt = (time.perf_counter() - t0) / 5
    print(chunk, f"{dt*1e3:.3f} ms  {n/dt:.3e} parcel/s  {40*n/dt/1e9:.1f} GB/s each way")
