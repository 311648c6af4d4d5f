#!/bin/bash
# time every variants/*.so on the same box: scripts/ab.sh [batch] [rounds]
B=${1:-262144}; R=${2:-2}
for r in $(seq 1 $R); do
  for f in variants/*.so; do
    echo "== $f"; C2A_B200_LIB=$PWD/$f timeout 300 python scripts/phase_stats.py $B 2>&1 | tail -5 | head -5
  done
done
