#!/bin/bash
# development sweep: lanes per world (LLE_B200_MIN_WD) x worlds per ticket (LLE_B200_GROUP) on small-agent-count levels
for lvl in 1 3; do
  for wd in 1 2 4; do
    for g in 8 16 32; do
      echo "lvl=$lvl MIN_WD=$wd GROUP=$g $(LLE_B200_MIN_WD=$wd LLE_B200_GROUP=$g python scratch/quick_bench.py --level $lvl --steps 3000 | cut -c80-200)"
    done
  done
done
