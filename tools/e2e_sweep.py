import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cloudy_b200 as cb
from cloudy_b200 import workloads as W
n = 1 << 20
par, state = W.c2_gamma_exp(n)
model = cb.CoalescenceModel(par)
hin = torch.from_numpy(state).pin_memory().numpy(); hout = torch.empty((n, 5), dtype=torch.float64).pin_memory().numpy()
for chunk in (56832, 113664, 131072, 170496, 227328, 262144, 349525):
    os.environ["CLOUDY_PIPE_CHUNK"] = str(chunk)
    for _ in range(3): model.coal_tendency_host(hin, hout)
    t0 = time.perf_counter()
    for _ in range(10): model.coal_tendency_host(hin, hout)
    dt = (time.perf_counter() - t0) / 10
    print(chunk, f"{dt*1e3:.3f} ms  {n/dt:.3e}")
