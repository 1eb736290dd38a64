#!/bin/bash
N=${1:-16}
for lib in build/lib_lov3.so build/lib_lov4.so build/lib_lov5.so; do
  for split in 0 200000 1000000; do
    SPLIT=$split CSOLVE_B200_LIB=$PWD/$lib python scripts/profile_target.py $N 2>&1 | sed "s|^|$lib split=$split |" | cut -c1-60,150-
  done
done
