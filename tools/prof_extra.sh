#!/bin/bash
# Profiling pass over the non-default configs (run under gpurun): launch lists + one full-set capture per kernel.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
python tools/bench_configs.py moving > $O/moving_before.jsonl 2>&1
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/launches_c3_r1b.csv python tools/bench_configs.py c3 > $O/c3_under_ncu.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 100 --csv --log-file $O/launches_moving_r1b.csv python tools/bench_configs.py moving > $O/moving_under_ncu.log 2>&1
$NCU --set full --import-source on -k regex:tpp_kernel -s 3 -c 1 -o $O/prof_r1b_moving -f python tools/bench_configs.py moving > $O/p1.log 2>&1
$NCU --set full --import-source on -k regex:flux_kernel -s 160 -c 1 -o $O/prof_r1b_flux -f python tools/bench_configs.py c3 > $O/p2.log 2>&1
$NCU --set full --import-source on -k regex:tpp_kernel -s 160 -c 1 -o $O/prof_r1b_c3tpp -f python tools/bench_configs.py c3 > $O/p3.log 2>&1
tail -3 $O/moving_before.jsonl
