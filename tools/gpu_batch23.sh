#!/bin/bash
O=gpurun_out
for lib in a b c d e; do
  echo $lib >> $O/r2v_ab.jsonl
  LLE_B200_LIB=$PWD/lle_b200/_native/liblle_b200_$lib.so python tools/bench_config.py --config 3 --repeat 2 >> $O/r2v_ab.jsonl 2>> $O/r2v_err.log
done
echo c8 >> $O/r2v_ab.jsonl
LLE_B200_TINY_E=8 LLE_B200_LIB=$PWD/lle_b200/_native/liblle_b200_c.so python tools/bench_config.py --config 3 --repeat 2 >> $O/r2v_ab.jsonl 2>> $O/r2v_err.log
echo a8 >> $O/r2v_ab.jsonl
LLE_B200_TINY_E=8 LLE_B200_LIB=$PWD/lle_b200/_native/liblle_b200_a.so python tools/bench_config.py --config 3 --repeat 2 >> $O/r2v_ab.jsonl 2>> $O/r2v_err.log
