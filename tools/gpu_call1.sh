#!/bin/bash
# round-2 call 1: sort/tile sweep (time + DRAM traffic) on C5, C5-share, C4
O=gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
python tools/sweep_sort.py c5 off whole 262144 524288 1048576 2097152 4194304 > $O/r02_sweep_c5.jsonl 2> $O/r02_sweep_c5.err
python tools/sweep_sort.py c4 off whole 524288 1048576 2097152 > $O/r02_sweep_c4.jsonl 2> $O/r02_sweep_c4.err
python tools/sweep_sort.py c2 off whole 262144 > $O/r02_sweep_c2.jsonl 2> $O/r02_sweep_c2.err
SWEEP_REPS=1 ncu --clock-control none -k regex:tpp_kernel --metrics $M --csv --log-file $O/r02_sweep_c5_ncu.csv python tools/sweep_sort.py c5 off whole 262144 524288 1048576 2097152 4194304 > $O/r02_sweep_c5_ncu.log 2>&1
SWEEP_REPS=1 ncu --clock-control none -k regex:tpp_kernel --metrics $M --csv --log-file $O/r02_sweep_c4_ncu.csv python tools/sweep_sort.py c4 off whole 524288 1048576 2097152 > $O/r02_sweep_c4_ncu.log 2>&1
cat $O/r02_sweep_c5.jsonl $O/r02_sweep_c4.jsonl $O/r02_sweep_c2.jsonl
