#!/bin/bash
O=gpurun_out
for lib in m7 m8 m9 m6; do
  echo $lib >> $O/r2x_ab.jsonl
  LLE_B200_LIB=$PWD/lle_b200/_native/liblle_b200_$lib.so python tools/bench_config.py --config 3 --repeat 2 >> $O/r2x_ab.jsonl 2>> $O/r2x_err.log
done
echo m7e8 >> $O/r2x_ab.jsonl
LLE_B200_TINY_E=8 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2x_ab.jsonl 2>> $O/r2x_err.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -x > $O/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2x_pytest.log
