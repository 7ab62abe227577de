#!/bin/bash
# Round-end measurement pass (run under gpurun): bench line, other configs, launch lists and one full-set capture per hot kernel.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
python bench.py > $O/bench_c2_final.json 2> $O/bench_c2_final.err
python tools/bench_configs.py c3 c4 c5 moving > $O/bench_configs_final.jsonl 2>&1
$NCU --metrics gpu__time_duration.sum -c 120 --csv --log-file $O/launches_final_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/l1.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/launches_final_c3.csv python tools/bench_configs.py c3 > $O/l2.log 2>&1
$NCU --set full --import-source on -k regex:tpp_kernel -s 5 -c 1 -o $O/prof_final_c2 -f python bench.py --no-cpu-baseline --steps 3 > $O/p1.log 2>&1
python tools/ncu_summary.py $O/prof_final_c2.ncu-rep 45 > $O/prof_final_c2_summary.txt 2>&1
$NCU --set full --import-source on -k regex:tpp_kernel -s 3 -c 1 -o $O/prof_final_moving -f python tools/bench_configs.py moving > $O/p2.log 2>&1
python tools/ncu_summary.py $O/prof_final_moving.ncu-rep 45 > $O/prof_final_moving_summary.txt 2>&1
rm -f $O/prof_final_moving.ncu-rep
$NCU --set full --import-source on -k regex:tpp_kernel -s 11 -c 1 -o $O/prof_final_moving4 -f python tools/bench_configs.py moving > $O/p3.log 2>&1
python tools/ncu_summary.py $O/prof_final_moving4.ncu-rep 45 > $O/prof_final_moving4_summary.txt 2>&1
rm -f $O/prof_final_moving4.ncu-rep
$NCU --set full --import-source on -k regex:flux_kernel -s 160 -c 1 -o $O/prof_final_flux -f python tools/bench_configs.py c3 > $O/p4.log 2>&1
python tools/ncu_summary.py $O/prof_final_flux.ncu-rep 45 > $O/prof_final_flux_summary.txt 2>&1
rm -f $O/prof_final_flux.ncu-rep
$NCU --set full --import-source on -k regex:tpp_kernel -s 160 -c 1 -o $O/prof_final_c3tpp -f python tools/bench_configs.py c3 > $O/p5.log 2>&1
python tools/ncu_summary.py $O/prof_final_c3tpp.ncu-rep 45 > $O/prof_final_c3tpp_summary.txt 2>&1
rm -f $O/prof_final_c3tpp.ncu-rep
$NCU --set full --import-source on -k regex:tpp_kernel -s 3 -c 1 -o $O/prof_final_c4 -f python tools/bench_configs.py c4 > $O/p6.log 2>&1
python tools/ncu_summary.py $O/prof_final_c4.ncu-rep 45 > $O/prof_final_c4_summary.txt 2>&1
rm -f $O/prof_final_c4.ncu-rep
$NCU --set full -k regex:regime_key -s 5 -c 1 -o $O/prof_final_key -f python bench.py --no-cpu-baseline --steps 3 > $O/p7.log 2>&1
python tools/ncu_summary.py $O/prof_final_key.ncu-rep 45 > $O/prof_final_key_summary.txt 2>&1
rm -f $O/prof_final_key.ncu-rep
cat $O/bench_c2_final.json | cut -c1-600
