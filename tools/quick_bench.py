"""Quick device-side timing of the coalescence tendency kernel (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
which = sys.argv[2] if len(sys.argv) > 2 else "c2"
if which == "c2":
    par, state = W.c2_gamma_exp(n)
elif which == "c4":
    par, state = W.c4_three_modes(n)
model = cb.CoalescenceModel(par)
ctx = model.ctx
print("fp64 peak TFLOP/s:", ctx.measure_fp64_peak())
u = model.ensemble(n).upload(state)
du = model.ensemble(n)
for lanes in (1, 4, 8):
    ctx.set_lanes(lanes)
    for _ in range(3):
        model.coal_tendency(u, du)
    ctx.sync()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        model.coal_tendency(u, du)
    ctx.sync()
    dt = (time.perf_counter() - t0) / reps
    print(f"{which} lanes={lanes:2d} n={n} {dt*1e3:.3f} ms  {n/dt:.3e} parcel-RHS/s")
