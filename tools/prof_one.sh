#!/bin/bash
# one full-set ncu capture of the thread-per-parcel kernel for a library variant: tools/prof_one.sh <tag> <config> [skip]
O=gpurun_out
tag=$1; cfgname=${2:-c2}; skip=${3:-2}
[ "$tag" != "main" ] && export CLOUDY_LIB=$PWD/cloudy.jl_b200/libcloudy_b200_$tag.so
SWEEP_REPS=1 ncu --set full --import-source on --clock-control none -k regex:tpp_kernel -s $skip -c 1 -o $O/prof_${tag}_$cfgname -f python tools/sweep_sort.py $cfgname resident > $O/prof_${tag}_$cfgname.log 2>&1
python tools/ncu_summary.py $O/prof_${tag}_$cfgname.ncu-rep 40 > $O/prof_${tag}_${cfgname}_summary.txt 2>&1
