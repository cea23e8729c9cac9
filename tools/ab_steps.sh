#!/bin/bash
# A/B timing of builds of libtdz.so on ONE box: tools/ab_steps.sh "STEP1 STEP2 ..." lib1.so lib2.so ...
steps=$1; shift
for rep in 1 2; do
  for lib in "$@"; do
    for s in $steps; do
      echo -n "$(basename $lib) "; TDZ_LIB=$PWD/$lib python tools/run_step.py $s 64 64000 5
    done
  done
done
