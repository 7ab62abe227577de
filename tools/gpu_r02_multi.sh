#!/bin/bash
# multi-GPU pass (run under gpurun --gpus N): tools/gpu_r02_multi.sh N  -> multirank test, PCIe probe with all ranks active, bench
N=${1:-2}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
nvidia-smi topo -m > $O/r02_topo_n$N.txt 2>&1
$TR --master-port 29533 tools/pcie_probe.py > $O/r02_pcie_probe_n$N.jsonl 2> $O/r02_pcie_probe_n$N.err
cat $O/r02_pcie_probe_n$N.jsonl | tail -4
$TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > $O/r02_bench_c5_n$N.json 2> $O/r02_bench_c5_n$N.err
python - <<PY
import json
d=json.load(open("$O/r02_bench_c5_n$N.json"))
print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","scaling","gpu_launches")}); print("e2e",d["e2e"]["value"],"resident",d.get("e2e_resident",{}).get("value"),"c2",d.get("c2_1mi_per_gpu",{}).get("value")); print(d["conservation"]); print(d["clocks"])
PY
tail -2 $O/r02_bench_c5_n$N.err
