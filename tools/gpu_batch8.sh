#!/bin/bash
O=gpurun_out
for c in 1 2 4 8 16; do
  LLE_B200_TINY_CHUNK=$c python tools/bench_config.py --config 3 --repeat 2 >> $O/r2h_cfg3.jsonl 2>> $O/r2h_err.log
done
LLE_B200_TINY_CHUNK=8 LLE_B200_TINY_E=4 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2h_cfg3.jsonl 2>> $O/r2h_err.log
LLE_B200_TINY_CHUNK=8 LLE_B200_TINY_E=16 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2h_cfg3.jsonl 2>> $O/r2h_err.log
for c in 1 2 4; do
  LLE_B200_CHUNK=$c python tools/bench_config.py --config 2 --repeat 2 >> $O/r2h_cfg2.jsonl 2>> $O/r2h_err.log
  LLE_B200_CHUNK=$c python tools/step_trace.py --config 2 >> $O/r2h_grid.jsonl 2>> $O/r2h_err.log
done
LLE_B200_NO_TINY=1 LLE_B200_CHUNK=4 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2h_cfg3.jsonl 2>> $O/r2h_err.log
LLE_B200_CHUNK=2 python tools/bench_config.py --config 1 --repeat 2 >> $O/r2h_cfg2.jsonl 2>> $O/r2h_err.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest.log
