#!/bin/bash
for g in 8 16 32; do echo "lvl6 GROUP=$g $(LLE_B200_GROUP=$g python scratch/quick_bench.py --level 6 --steps 4000 | cut -c80-200)"; done
for tt in 1024 2048 4096; do for g in 16 32; do
  echo "gen TILE_TARGET=$tt GROUP=$g $(LLE_B200_TILE_TARGET=$tt LLE_B200_GROUP=$g python scratch/quick_bench.py --map generated --envs 1048576 --steps 400 | cut -c90-200)"; done; done
for wd in 1 2; do echo "gen MIN_WD=$wd $(LLE_B200_MIN_WD=$wd python scratch/quick_bench.py --map generated --envs 1048576 --steps 400 | cut -c90-200)"; done
for l in 2 4 5; do echo "lvl$l $(python scratch/quick_bench.py --level $l --steps 3000 | cut -c80-200)"; done
