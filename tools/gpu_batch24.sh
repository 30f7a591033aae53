#!/bin/bash
O=gpurun_out
for lib in f g h i j d; do
  echo $lib >> $O/r2w_ab.jsonl
  LLE_B200_LIB=$PWD/lle_b200/_native/liblle_b200_$lib.so python tools/bench_config.py --config 3 --repeat 2 >> $O/r2w_ab.jsonl 2>> $O/r2w_err.log
done
echo h8 >> $O/r2w_ab.jsonl
LLE_B200_TINY_E=8 LLE_B200_LIB=$PWD/lle_b200/_native/liblle_b200_h.so python tools/bench_config.py --config 3 --repeat 2 >> $O/r2w_ab.jsonl 2>> $O/r2w_err.log
