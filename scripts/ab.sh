#!/bin/bash
# A/B of builds of the library on the same box: ab.sh "<lib A> <lib B> ..." [probe args]
LIBS=$1; shift
for rep in 1 2; do
  for L in $LIBS; do
    CSOLVE_B200_LIB=$L python scripts/probe.py "$@" 2>&1 | sed "s|^|$(basename $L) |" | cut -c1-50,120-250
  done
done
