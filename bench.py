#!/usr/bin/env python
"""bench.py — coal-tendency evals/s (parcels x RHS) on N B200s, with roofline and CPU baseline.

  python bench.py --gpus 1 --steps 20 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the reference's CPU path (oracle restatement) on the host cores

A "step" is one evaluation of the coalescence right-hand side (rhs_coal!, box_model_helpers.jl:29-53) over every
parcel of the rank's ensemble: BASELINE.json configs[1] — Gamma cloud + Exponential rain modes, linear (Golovin)
kernel, 1,048,576 synthetic parcels per GPU (weak scaling: parcels shard with no exchange; every
--sums-every steps the per-slot moment sums are all-reduced with NCCL, the path's only collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "coal-tendency evals/sec (parcels x RHS)"
UNIT = "parcel-RHS/s"
N_PARCELS = 1 << 20
# nominal algorithmic work per parcel-RHS for this workload (restated in DESIGN.md):
FLOP_PER_EVAL = 1.96e4   # 75 nodes x 256 flop + ~0.4k contraction
BYTES_PER_EVAL = 80.0    # read 5 + write 5 doubles


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--parcels", type=int, default=N_PARCELS, help="parcels per GPU")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--sums-every", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload(n, seed_offset=0):
    from cloudy_b200 import workloads as W
    return W.c2_gamma_exp(n_parcels=n, seed=W.SEED0 + 2 + seed_offset)


def config_dict(args, world):
    return {
        "workload": "C2 box model: Gamma cloud + Exponential rain, linear (Golovin) kernel, NProgMoms=(3,2) "
                    "(5 moments, 6-slot reading of BASELINE.json), thresholds (0.5, Inf) normalised",
        "parcels_per_gpu": args.parcels,
        "global_parcels": args.parcels * world,
        "parallelism": f"parcels block-partitioned over {world} GPU(s), no halo; NCCL all-reduce of 5 moment sums every "
                       f"{args.sums_every} steps" if world > 1 else "single GPU",
        "l2": "4 rotating input/output ensemble pairs (>= 336 MB at 1Mi parcels) so no step re-reads L2-resident data",
        "step": "one rhs_coal! evaluation over every parcel of the rank",
    }


# ---------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle's structure-faithful C/OpenMP restatement on the host cores
# ---------------------------------------------------------------------------------------------------
def host_threads():
    """all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the scheduler instead)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_rate(par, state, seconds, threads):
    import cloudy_b200 as cb
    from oracle import c_oracle
    cfg = cb.build_config(tuple(d.kind for d in par.pdists), par.coal_data, norms=par.norms)
    probe = min(state.shape[0], 256 * threads)
    t0 = time.perf_counter()
    c_oracle.rhs_coal_batch(cfg, state[:probe], n_threads=threads)
    r0 = probe / (time.perf_counter() - t0)
    n = int(max(probe, min(state.shape[0], r0 * seconds)))
    t0 = time.perf_counter()
    c_oracle.rhs_coal_batch(cfg, state[:n], n_threads=threads)
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    threads = host_threads()
    sample = 2048 * threads
    par, state = workload(sample)
    import cloudy_b200 as cb
    cfg = cb.build_config(tuple(d.kind for d in par.pdists), par.coal_data, norms=par.norms)
    for _ in range(args.warmup):
        c_oracle.rhs_coal_batch(cfg, state[: 256 * threads], n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.rhs_coal_batch(cfg, state, n_threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(config_dict(args, 1), step=f"one rhs_coal! evaluation over a bounded sample of {sample} parcels of the same workload"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} parcels x {args.steps} steps; C/OpenMP structure-faithful restatement of the Julia path "
                                   "(Julia is not installed; oracle/cloudy_oracle.c)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Keep the rank's threads (and therefore its first-touch pinned host buffers) on the CPU cores
    closest to its GPU, so that the host<->device copies of the e2e leg do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def run_b200(args):
    import torch
    import torch.distributed as dist
    import cloudy_b200 as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a non-default torch stream is made current for the whole run and handed to the library, so that
    # torch.cuda.Event timing brackets the library's launches (stream handle 0 would mean "private stream")
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    ctx = cb.Context(local, stream=stream)
    n = args.parcels
    par, state0 = workload(n, seed_offset=1000 * rank)
    model = cb.CoalescenceModel(par, ctx=ctx)
    if args.lanes:
        ctx.set_lanes(args.lanes)
    NBUF = 4
    ins, outs = [], []
    for b in range(NBUF):
        st = state0 if b == 0 else np.roll(state0, b * 977, axis=0)
        ins.append(model.ensemble(n).upload(st))
        outs.append(model.ensemble(n))
    sums = torch.zeros(model.n_slots, dtype=torch.float64, device="cuda")
    fp64_peak = ctx.measure_fp64_peak()

    def step(i):
        b = i % NBUF
        model.coal_tendency(ins[b], outs[b])
        if world > 1 and (i + 1) % args.sums_every == 0:
            model.moment_sums_device(outs[b], sums.data_ptr())
            dist.all_reduce(sums)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        kev[i][0].record()
        model.coal_tendency(ins[i % NBUF], outs[i % NBUF])
        kev[i][1].record()
        if world > 1 and (i + 1) % args.sums_every == 0:
            model.moment_sums_device(outs[i % NBUF], sums.data_ptr())
            dist.all_reduce(sums)
    t_end.record()
    barrier()
    launches = ctx.launch_count() - l0
    total_ms = t_start.elapsed_time(t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in kev]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    # ---- sustained window: the same step repeated for ~1 s with the clock sampler running (the K timed steps above last
    # only milliseconds, too short for nvidia-smi's sampling period); reported next to the burst value
    sus_steps = int(min(5000, max(50, 1000.0 / max(total_ms / args.steps, 1e-3))))
    s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(sus_steps):
        model.coal_tendency(ins[i % NBUF], outs[i % NBUF])
    s1.record()
    barrier()
    ts_ = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
    sustained_value = n * world * sus_steps / (float(ts_.item()) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = f"timed region + {sus_steps} further identical steps ({float(ts_.item()):.0f} ms)"
    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + kernel + D2H per step) ----
    h_in = torch.from_numpy(state0).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    hin_np, hout_np = h_in.numpy(), h_out.numpy()
    e2e_steps = max(3, args.steps)
    for _ in range(3):
        model.coal_tendency_host(hin_np, hout_np)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        model.coal_tendency_host(hin_np, hout_np)
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n * world * e2e_steps / (float(te.item()) * 1e-3)
    checksum = float(np.abs(hout_np[:1024]).sum())

    if rank == 0:
        value = n * world * args.steps / (total_ms * 1e-3)
        k_ms = float(np.mean(kernel_ms))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        achieved_tf = FLOP_PER_EVAL * n / (k_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(args, world),
            "pair_evals_per_s": value * 4,  # parcel-mode-pair evals/s = value x N^2
            "sustained": {"value": sustained_value, "unit": UNIT, "steps": sus_steps},
            "roofline": {
                "bound": "fp64", "kernel": "tpp_kernel<2,2,BOX>", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak if fp64_peak else None,
                "peak_source": "DFMA-chain microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry); nominal 37.2",
                "flop_per_eval": FLOP_PER_EVAL, "kernel_ms": k_ms,
                # sm__pipe_fp64_cycles_active of the same kernel and workload, ncu --set full (not measured in this run):
                "fp64_pipe_active_ncu": 0.535 if n == (1 << 20) else None,
                "note": "achieved = nominal algorithmic flop model of SURVEY 8(d) (two exp + one incomplete gamma per node and integral "
                        "entry sharing none) x parcels / time; the kernel EXECUTES ~4x fewer FP64 instructions than that model (one exp per "
                        "node, Taylor evaluation of the shared incomplete gamma), so frac ~ 1 coexists with ~50% FP64-pipe utilisation "
                        "(profiles/r01_tpp_kernel_c2_ncu_full_summary.txt); kernel_ms includes the 2 regime-sort launches of each step",
                "hbm": {"achieved": BYTES_PER_EVAL * n / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": BYTES_PER_EVAL * n / (k_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
                # dram__bytes_read.sum + dram__bytes_write.sum of one tpp_kernel launch over 1 Mi parcels, ncu --set full
                # (profiles/r01_tpp_kernel_c2_ncu_full_summary.txt): 84.5 + 39.1 MB vs 84 MB algorithmic (the regime-sorted gather
                # touches 32-byte sectors for 8-byte loads); irrelevant to the FP64-bound duration
                "traffic": 123.6e6 if n == (1 << 20) else None,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(state0.nbytes), "d2h_bytes_per_step": int(state0.nbytes),
                    "steps": e2e_steps, "api": "cloudy_coal_tendency_host (pinned host buffers)", "checksum": checksum},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core again
            threads = host_threads()
            rate, ns, dt = cpu_rate(par, state0, args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"first {ns} parcels of the same ensemble, {dt:.1f} s; C/OpenMP structure-faithful "
                                              "restatement of the Julia path (oracle/cloudy_oracle.c)"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def ensure_built():
    """the native libraries are build products: compile them if this checkout has none (nvcc / gcc are in the image)"""
    if not (os.path.exists(os.path.join(ROOT, "cloudy.jl_b200", "libcloudy_b200.so")) and
            os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so"))):
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__
            __graft_entry__.build()
        else:
            while not os.path.exists(os.path.join(ROOT, "cloudy.jl_b200", "libcloudy_b200.so")):
                time.sleep(1.0)
            time.sleep(2.0)


_REAL_STDOUT = None


def emit(line):
    """the one JSON line of the contract, on the process's real stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    # Libraries may write to fd 1 (NCCL prints its version banner there when NCCL_DEBUG is set): everything except the
    # result line goes to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ensure_built()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
