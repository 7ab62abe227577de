#!/bin/bash
# round-2 call 2: full GPU test-suite on the resident-order build, sort sweep (time + DRAM traffic)
O=gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r02_call2_pytest.log
SWEEP_STEP=1 python tools/sweep_sort.py c5 off resident > $O/r02_sweep2_c5.jsonl 2> $O/r02_sweep2_c5.err
python tools/sweep_sort.py c4 off resident > $O/r02_sweep2_c4.jsonl 2> $O/r02_sweep2_c4.err
python tools/sweep_sort.py c2 off resident > $O/r02_sweep2_c2.jsonl 2> $O/r02_sweep2_c2.err
python tools/sweep_sort.py c5s off resident > $O/r02_sweep2_c5s.jsonl 2> $O/r02_sweep2_c5s.err
SWEEP_REPS=1 ncu --clock-control none -k regex:"tpp_kernel|sort_|regime_key" --metrics $M --csv --log-file $O/r02_sweep2_c5_ncu.csv python tools/sweep_sort.py c5 resident > $O/r02_sweep2_c5_ncu.log 2>&1
SWEEP_REPS=1 ncu --clock-control none -k regex:"tpp_kernel|sort_|regime_key" --metrics $M --csv --log-file $O/r02_sweep2_c4_ncu.csv python tools/sweep_sort.py c4 resident > $O/r02_sweep2_c4_ncu.log 2>&1
cat $O/r02_call2_pytest.log $O/r02_sweep2_*.jsonl
tail -3 $O/r02_sweep2_*.err
