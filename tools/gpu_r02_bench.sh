#!/bin/bash
# round-2 measurement pass (run under gpurun): tools/gpu_r02_bench.sh <git-hash>
# bench line on C5, launch list of the same command, full-set captures of the hot kernel on 64 Mi and 8 Mi parcels
O=gpurun_out
GIT=${1:-unknown}
NCU="ncu --clock-control none"
python bench.py > $O/r02_bench_c5_n1.json 2> $O/r02_bench_c5_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $O/r02_launches_bench_c5.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/r02_l1.log 2>&1
python tools/launch_summary.py $O/r02_launches_bench_c5.csv > $O/r02_launches_bench_c5_summary.csv 2>&1
for c in c5 c5s; do
  [ $c = c5 ] && P=67108864 || P=8388608
  SWEEP_REPS=1 $NCU --set full --import-source on -k regex:tpp_kernel -s 1 -c 1 -o $O/r02_prof_$c -f python tools/sweep_sort.py $c resident > $O/r02_prof_$c.log 2>&1
  python tools/ncu_summary.py $O/r02_prof_$c.ncu-rep 40 --json $O/r02_tpp_kernel_${c}_ncu_full.json parcels=$P git=$GIT config=$c > $O/r02_tpp_kernel_${c}_ncu_full_summary.txt 2>&1
done
rm -f $O/r02_prof_c5s.ncu-rep
cut -c1-1500 $O/r02_bench_c5_n1.json; tail -3 $O/r02_bench_c5_n1.err
cut -c1-300 $O/r02_bench_ref.json
cat $O/r02_launches_bench_c5_summary.csv | head -20
