#!/bin/bash
# development sweep: occupancy (LLE_MIN_CTAS, compile time) x tile buffers (LLE_B200_NBUF) on three shapes
for ctas in 4 5 6; do
  LLE_B200_NVCC_FLAGS="-DLLE_MIN_CTAS=$ctas" python lle_b200/build.py --force --ptxas 2>&1 | grep -E "spill" | sort | uniq -c | head -3
  for nb in 1 2; do
    echo "== MIN_CTAS=$ctas NBUF=$nb"
    export LLE_B200_NBUF=$nb
    python scratch/quick_bench.py --level 1 --steps 3000
    python scratch/quick_bench.py --level 6 --steps 3000
    python scratch/quick_bench.py --map generated --envs 1048576 --steps 400
  done
done
