#!/bin/bash
O=gpurun_out
for lib in liblle_b200_v1.so liblle_b200_v2.so liblle_b200_v3.so liblle_b200.so; do
  echo $lib >> $O/r2u_ab.jsonl
  LLE_B200_TINY_E=4 LLE_B200_LIB=$PWD/lle_b200/_native/$lib python tools/bench_config.py --config 3 --repeat 2 >> $O/r2u_ab.jsonl 2>> $O/r2u_err.log
done
