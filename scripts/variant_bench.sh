#!/bin/bash
N=${1:-16}
for split in 0 1000000 3000000 8000000; do
  for rep in 1 2; do
    SPLIT=$split python scripts/profile_target.py $N 2>&1 | sed "s|^|split=$split |" | cut -c1-40,120-
  done
done
