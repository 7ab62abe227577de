"""Regime-sort sweep on one GPU: time (CUDA events) of one tendency evaluation for each sort variant.
   python tools/sweep_sort.py c5 off whole 262144 1048576 ...   (variants: off | whole | <tile size in parcels>)
Run it under `ncu -k regex:tpp_kernel --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
to get the DRAM traffic of the same launches (the script prints the launch order)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cloudy_b200 as cb
from cloudy_b200 import workloads as W

cfg = sys.argv[1]
variants = sys.argv[2:] or ["off", "whole"]
reps = int(os.environ.get("SWEEP_REPS", "3"))
torch.cuda.set_device(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
ctx = cb.Context(0, stream=ts.cuda_stream)
if cfg == "c5":
    n = 1 << 26
    par, base = W.c2_gamma_exp(1 << 23)
    state = np.tile(base, (8, 1))
elif cfg == "c5s":  # one GPU's share of C5 at 8 GPUs
    n = 1 << 23
    par, state = W.c2_gamma_exp(n)
elif cfg == "c2":
    n = 1 << 20
    par, state = W.c2_gamma_exp(n)
elif cfg == "c4":
    n = 1 << 24
    par, base = W.c4_three_modes(1 << 22)
    state = np.tile(base, (4, 1))
else:
    raise SystemExit("unknown config")
model = cb.CoalescenceModel(par, ctx=ctx)
u = model.ensemble(n); du = model.ensemble(n)
def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for v in variants:
    u.upload(state)
    ctx.set_regime_sort(0 if v == "off" else 2)
    sort_ms = None
    if v != "off":
        sort_ms = timed(lambda: u.regime_sort(), 1)      # first sort of the host order
        resort_ms = timed(lambda: u.regime_sort(), reps)  # refresh of an already ordered ensemble
    model.coal_tendency(u, du)
    ms = timed(lambda: model.coal_tendency(u, du), reps)
    step_ms = timed(lambda: model.ssprk33_steps(u, 1e-3, 1, cb.MODEL_BOX), reps) if os.environ.get("SWEEP_STEP") else None
    print(json.dumps({"config": cfg, "variant": v, "parcels": n, "ms": ms, "parcel_rhs_per_s": n / ms * 1e3, "sort_ms": sort_ms,
                      "resort_ms": resort_ms if v != "off" else None, "ssprk33_step_ms": step_ms, "tpp_launches": reps + 1}), flush=True)
