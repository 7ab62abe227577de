"""Documentation runs of the other BASELINE.json configs on one GPU (bench.py, the driver's contract, measures configs[1]):
   C3 rainshaft 4096 columns x 256 levels (RHS and fused SSPRK33 step), C4 order-4 kernel / 3 modes at 16 Mi parcels,
   C5's ensemble (C2 generator) at 64 Mi parcels.  One JSON line per config; CUDA-event timing, >= 3 warm-ups."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cloudy_b200 as cb
from cloudy_b200 import workloads as W

torch.cuda.set_device(0)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
ctx = cb.Context(0, stream=ts.cuda_stream)
peak = ctx.measure_fp64_peak()


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


which = sys.argv[1:] or ["c3", "c4", "c5"]
if "c3" in which:
    par, cols = W.c3_rainshaft(4096, 256)
    state = cols.reshape(-1, 6); n = state.shape[0]
    model = cb.CoalescenceModel(par, ctx=ctx, nz=256)
    u = model.ensemble(n).upload(state); du = model.ensemble(n)
    model.ssprk33_steps(u, par.dt, 50, cb.MODEL_RAINSHAFT)  # let rain form and fall so that more cells are active
    active = float((u.download()[:, 0] > 0).mean())
    ms_rhs = timed(lambda: model.rainshaft_rhs(u, du), 10)
    ms_step = timed(lambda: model.ssprk33_steps(u, par.dt, 1, cb.MODEL_RAINSHAFT), 10)
    print(json.dumps({"config": "C3 rainshaft 4096 columns x 256 levels, 2 Gamma modes, after 50 steps", "cells": n, "active_cell_fraction": active,
                      "rhs_ms": ms_rhs, "cell_rhs_per_s": n / ms_rhs * 1e3, "ssprk33_step_ms": ms_step, "cell_steps_per_s": n / ms_step * 1e3,
                      "hbm_GBps_step": 384.0 * n / ms_step * 1e3 / 1e9}))
    del u, du, model
if "c4" in which:
    n = 1 << 24
    par, state = W.c4_three_modes(n)
    model = cb.CoalescenceModel(par, ctx=ctx)
    u = model.ensemble(n).upload(state); du = model.ensemble(n)
    ms = timed(lambda: model.coal_tendency(u, du), 3)
    print(json.dumps({"config": "C4 3 Gamma modes, order-4 hydrodynamic tensor (P=5), thresholds (1,100,Inf), 16 Mi parcels", "parcels": n,
                      "ms": ms, "parcel_rhs_per_s": n / ms * 1e3, "pair_evals_per_s": 9 * n / ms * 1e3,
                      "nominal_tflops": 6.2e4 * n / ms * 1e3 / 1e12, "fp64_peak_tflops": peak}))
    del u, du, model, state
if "c5" in which:
    n = 1 << 26
    par, state = W.c2_gamma_exp(n)
    model = cb.CoalescenceModel(par, ctx=ctx)
    u = model.ensemble(n).upload(state); du = model.ensemble(n)
    del state
    ms = timed(lambda: model.coal_tendency(u, du), 3)
    ms_step = timed(lambda: model.ssprk33_steps(u, 1.0, 1, cb.MODEL_BOX), 3)
    print(json.dumps({"config": "C5 ensemble on ONE GPU: Gamma+Exponential, 64 Mi parcels", "parcels": n, "rhs_ms": ms,
                      "parcel_rhs_per_s": n / ms * 1e3, "ssprk33_step_ms": ms_step, "parcel_steps_per_s": n / ms_step * 1e3,
                      "nominal_tflops": 1.96e4 * n / ms * 1e3 / 1e12, "fp64_peak_tflops": peak}))
if "moving" in which:
    n = 1 << 20
    par, state = W.moving_gamma_exp(n)
    model = cb.CoalescenceModel(par, ctx=ctx)
    u = model.ensemble(n).upload(state); du = model.ensemble(n)
    ms = timed(lambda: model.coal_tendency(u, du), 5)
    print(json.dumps({"config": "MovingThreshold: Gamma+Exponential (C2 ensemble), percentile 0.97, 1 Mi parcels", "parcels": n, "ms": ms,
                      "parcel_rhs_per_s": n / ms * 1e3}))
    par, state = W.moving_four_modes(n)
    model = cb.CoalescenceModel(par, ctx=ctx)
    u = model.ensemble(n).upload(state); du = model.ensemble(n)
    ms = timed(lambda: model.coal_tendency(u, du), 3)
    print(json.dumps({"config": "MovingThreshold: 4 Gamma modes, percentiles (0.99,0.99,0.99,1) (box_gamma_mix_moving.jl), 1 Mi parcels", "parcels": n,
                      "ms": ms, "parcel_rhs_per_s": n / ms * 1e3, "pair_evals_per_s": 16 * n / ms * 1e3}))
