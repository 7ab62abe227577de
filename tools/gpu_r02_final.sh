#!/bin/bash
# round-2 final measurement pass (run under gpurun): tools/gpu_r02_final.sh <git-hash>
O=gpurun_out
GIT=${1:-unknown}
NCU="ncu --clock-control none"
tools/gpu_r02_bench.sh $GIT > $O/r02_final_bench.log 2>&1
python tools/bench_configs.py c3 c4 moving > $O/r02_bench_other_configs_n1.jsonl 2> $O/r02_bench_other.err
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/r02_launches_c3_rainshaft.csv python tools/bench_configs.py c3 > $O/r02_l3.log 2>&1
python tools/launch_summary.py $O/r02_launches_c3_rainshaft.csv > $O/r02_launches_c3_rainshaft_summary.csv 2>&1
SWEEP_REPS=1 $NCU --set full --import-source on -k regex:tpp_kernel -s 1 -c 1 -o $O/r02_prof_c4 -f python tools/sweep_sort.py c4 resident > $O/r02_prof_c4.log 2>&1
python tools/ncu_summary.py $O/r02_prof_c4.ncu-rep 40 --json $O/r02_tpp_kernel_c4_ncu_full.json parcels=16777216 git=$GIT config=c4 > $O/r02_tpp_kernel_c4_ncu_full_summary.txt 2>&1
rm -f $O/r02_prof_c4.ncu-rep
$NCU --set full --import-source on -k regex:tpp_kernel -s 160 -c 1 -o $O/r02_prof_c3 -f python tools/bench_configs.py c3 > $O/r02_prof_c3.log 2>&1
python tools/ncu_summary.py $O/r02_prof_c3.ncu-rep 40 --json $O/r02_tpp_kernel_c3_rainshaft_ncu_full.json cells=1048576 git=$GIT config=c3 > $O/r02_tpp_kernel_c3_rainshaft_ncu_full_summary.txt 2>&1
rm -f $O/r02_prof_c3.ncu-rep
$NCU --set full -k regex:flux_kernel -s 160 -c 1 -o $O/r02_prof_flux -f python tools/bench_configs.py c3 > $O/r02_prof_flux.log 2>&1
python tools/ncu_summary.py $O/r02_prof_flux.ncu-rep 10 > $O/r02_flux_kernel_c3_ncu_full_summary.txt 2>&1
rm -f $O/r02_prof_flux.ncu-rep
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_c2_gamma_exp or test_c3_rainshaft_rhs or ragged" > $O/r02_compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r02_compute_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged" > $O/r02_compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r02_compute_sanitizer_racecheck.log
tail -12 $O/r02_final_bench.log | cut -c1-400
cat $O/r02_bench_other_configs_n1.jsonl | cut -c1-330
cat $O/r02_launches_c3_rainshaft_summary.csv
grep "duration\|fp64_cycles\|issue_active" $O/r02_tpp_kernel_c4_ncu_full_summary.txt $O/r02_tpp_kernel_c3_rainshaft_ncu_full_summary.txt $O/r02_tpp_kernel_c5_ncu_full_summary.txt $O/r02_flux_kernel_c3_ncu_full_summary.txt
tail -4 $O/r02_compute_sanitizer_memcheck.log
tail -4 $O/r02_compute_sanitizer_racecheck.log
