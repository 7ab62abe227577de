"""PCIe probe: pinned-host <-> device copy bandwidth, one direction and both at once (context for the e2e leg of bench.py).
Under torchrun every rank drives its own GPU AT THE SAME TIME (barrier before each timed section): the per-rank numbers show
what the host side (memory, root complex) gives each GPU when all of them copy together.
   python tools/pcie_probe.py
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe.py"""
import json, os
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1); h_out.fill_(2)   # first touch by this rank
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def numa_of_gpu(index):
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        return int(open(path).read().strip())
    except Exception:
        return None


for _ in range(2):
    run(True, True, 2)
res = {"rank": rank, "world": world, "gpu": local, "gpu_numa_node": numa_of_gpu(local), "cpus_allowed": len(os.sched_getaffinity(0)),
       "h2d_only_GBps": run(True, False), "d2h_only_GBps": run(False, True), "both_GBps_each": run(True, True), "bytes": n}
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        for r in out:
            print(json.dumps(r))
        print(json.dumps({"world": world, "sum_h2d_only_GBps": sum(r["h2d_only_GBps"] for r in out), "sum_d2h_only_GBps": sum(r["d2h_only_GBps"] for r in out),
                          "sum_both_GBps_each": sum(r["both_GBps_each"] for r in out)}))
    dist.destroy_process_group()
else:
    print(json.dumps(res))
