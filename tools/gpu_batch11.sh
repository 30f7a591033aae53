#!/bin/bash
O=gpurun_out
for e in 0 8 16 32; do
  LLE_B200_TINY_E=$e python tools/bench_config.py --config 3 --repeat 2 >> $O/r2k_cfg3.jsonl 2>> $O/r2k_err.log
done
LLE_B200_TINY_E=8 LLE_B200_TINY_CHUNK=2 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2k_cfg3.jsonl 2>> $O/r2k_err.log
LLE_B200_TINY_E=8 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -k "config3 or tiny or generated or heterogeneous or small or ragged or many" > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2k_pytest.log
CMD="python tools/bench_config.py --config 3 --steps 12 --warmup 4"
LLE_B200_TINY_E=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:lle_tiny_step_kernel -s 8 -c 1 -f -o $O/prof_cfg3_tiny7_r02 $CMD > $O/r2k_ncu_cfg3.log 2>&1
