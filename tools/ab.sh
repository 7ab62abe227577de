#!/bin/bash
# A/B timing of development library variants (cloudy.jl_b200/build.py, CLOUDY_DEV=<tag>): tools/ab.sh <tag> [<tag> ...]
O=gpurun_out
for v in "$@"; do
  export CLOUDY_LIB=$PWD/cloudy.jl_b200/libcloudy_b200_$v.so
  for c in ${AB_CONFIGS:-c2 c5s}; do
    python tools/sweep_sort.py $c resident 2> $O/ab_${v}_$c.err | sed "s/^/$v /"
  done
done
