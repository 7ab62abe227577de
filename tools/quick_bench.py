"""Quick device-side timing of the hot kernels (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cloudy_b200 as cb
from cloudy_b200 import workloads as W

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
lanes_list = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 8]
if which == "c2":
    par, state = W.c2_gamma_exp(n); nz = 1
elif which == "c4":
    par, state = W.c4_three_modes(n); nz = 1
elif which == "c3":
    nz = 256
    par, cols = W.c3_rainshaft(n // nz, nz); state = cols.reshape(-1, 6); n = state.shape[0]
model = cb.CoalescenceModel(par, nz=nz)
ctx = model.ctx
print("fp64 peak TFLOP/s:", ctx.measure_fp64_peak())
u = model.ensemble(n).upload(state)
du = model.ensemble(n)
sort = len(sys.argv) > 4 and sys.argv[4] == "sort"
ctx.set_regime_sort(1 if sort else 0)
for lanes in lanes_list:
    ctx.set_lanes(lanes)
    def run():
        if which == "c3":
            model.rainshaft_rhs(u, du)
        else:
            model.coal_tendency(u, du)
    for _ in range(3):
        run()
    ctx.sync()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        run()
    ctx.sync()
    dt = (time.perf_counter() - t0) / reps
    print(f"{which} lanes={lanes:2d} n={n} {dt*1e3:.3f} ms  {n/dt:.3e} evals/s")
if which == "c3":
    ctx.set_lanes(0)
    for _ in range(2):
        model.ssprk33_steps(u, 1.0, 2, cb.MODEL_RAINSHAFT)
    ctx.sync(); t0 = time.perf_counter()
    model.ssprk33_steps(u, 1.0, 10, cb.MODEL_RAINSHAFT); ctx.sync()
    dt = (time.perf_counter() - t0) / 10
    print(f"c3 ssprk33 step {dt*1e3:.3f} ms  {n/dt:.3e} cell-steps/s")
