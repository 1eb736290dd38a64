#!/bin/bash
# development helper: time queens N with several library builds
N=${1:-15}
for lib in build/lib_mb2.so build/lib_mb3.so build/lib_mb4.so; do
  for rep in 1 2; do
    CSOLVE_B200_LIB=$PWD/$lib python scripts/profile_target.py $N 2>&1 | sed "s|^|$lib |"
  done
done
