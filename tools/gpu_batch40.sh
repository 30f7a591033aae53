#!/bin/bash
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_c_client.py -m gpu -q --timeout 200 --timeout-method=thread -p no:cacheprovider -k "run_host_policy or quickstart or parts" > $O/r3n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r3n_pytest.log
for lib in "" _list2 "" _list2; do
  echo "lib$lib" >> $O/r3n_ab.jsonl
  LLE_B200_LIB=$PWD/lle_b200/_native/liblle_b200$lib.so python tools/bench_config.py --config 3 --repeat 2 >> $O/r3n_ab.jsonl 2>> $O/r3n_err.log
done
