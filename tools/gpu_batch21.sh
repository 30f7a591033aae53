#!/bin/bash
O=gpurun_out
for lib in liblle_b200_15077ed.so liblle_b200_head.so liblle_b200.so; do
  echo $lib >> $O/r2t_ab.jsonl
  LLE_B200_LIB=$PWD/lle_b200/_native/$lib python tools/bench_config.py --config 3 --repeat 2 >> $O/r2t_ab.jsonl 2>> $O/r2t_err.log
done
