#!/bin/bash
O=gpurun_out
run() { echo "$1" >> $O/r3o_ab.jsonl; shift; env "$@" python tools/bench_config.py --config 3 --repeat 2 >> $O/r3o_ab.jsonl 2>> $O/r3o_err.log; }
run "default(list2)" X=1
run "list3" LLE_B200_LIB=$PWD/lle_b200/_native/liblle_b200_list3.so
run "nbuf2 E=2" LLE_B200_TINY_NBUF=2 LLE_B200_TINY_E=2
run "nbuf2 E=4" LLE_B200_TINY_NBUF=2 LLE_B200_TINY_E=4
run "nbuf2 E=1" LLE_B200_TINY_NBUF=2 LLE_B200_TINY_E=1
run "default(list2)" X=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 --timeout-method=thread -p no:cacheprovider -k "config3 or levels or corpus" > $O/r3o_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r3o_pytest.log
LLE_B200_TINY_NBUF=2 LLE_B200_TINY_E=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 --timeout-method=thread -p no:cacheprovider -k "config3 or levels or corpus" > $O/r3o_pytest2.log 2>&1; echo "pytest rc=$?" >> $O/r3o_pytest2.log
