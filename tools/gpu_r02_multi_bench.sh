#!/bin/bash
# multi-GPU bench line only (run under gpurun --gpus N): tools/gpu_r02_multi_bench.sh N [test]
N=${1:-2}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
[ "$2" = "test" ] && python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
$TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > $O/r02_bench_c5_n$N.json 2> $O/r02_bench_c5_n$N.err
python - <<PY
import json
d=json.load(open("$O/r02_bench_c5_n$N.json"))
print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","scaling","gpu_launches")}); print("e2e",d["e2e"]["value"],"resident",d.get("e2e_resident",{}).get("value"),"c2",d.get("c2_1mi_per_gpu",{}).get("value")); print(d["conservation"]); print(d["clocks"])
PY
tail -2 $O/r02_bench_c5_n$N.err
