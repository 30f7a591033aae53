#!/bin/bash
O=gpurun_out
for sk in 0 1 3; do
  LLE_B200_TINY_SKIP=$sk python tools/bench_config.py --config 3 --repeat 2 >> $O/r2j_cfg3.jsonl 2>> $O/r2j_err.log
done
LLE_B200_TINY_SKIP=3 python tools/bench_config.py --config 3 --envs 262144 --repeat 2 >> $O/r2j_cfg3.jsonl 2>> $O/r2j_err.log
python tools/bench_config.py --config 3 --envs 262144 --repeat 2 >> $O/r2j_cfg3.jsonl 2>> $O/r2j_err.log
