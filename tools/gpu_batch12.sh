#!/bin/bash
O=gpurun_out
for e in 4 8; do
  LLE_B200_TINY_E=$e python tools/bench_config.py --config 3 --repeat 2 >> $O/r2l_cfg3.jsonl 2>> $O/r2l_err.log
done
LLE_B200_TINY_E=4 LLE_B200_TINY_CTAS_PER_SM=4 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2l_cfg3.jsonl 2>> $O/r2l_err.log
