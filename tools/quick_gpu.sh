#!/bin/bash
# quick regression + timing pass (run under gpurun)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/bench_configs.py ${1:-moving c3 c4} 2>&1 | tail -5 | cut -c1-330
for i in 1 2; do python bench.py --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2 value', d['value'], 'sustained', d['sustained']['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], 'clk', d['clocks']['sm_mhz'])"; done
