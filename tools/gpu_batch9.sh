#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_c_client.py -m gpu -q -p no:cacheprovider > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2i_pytest.log
python tools/bench_config.py --config 3 --repeat 2 >> $O/r2i_cfg3.jsonl 2>> $O/r2i_err.log
LLE_B200_TINY_CTAS_PER_SM=4 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2i_cfg3.jsonl 2>> $O/r2i_err.log
LLE_B200_TINY_CTAS_PER_SM=5 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2i_cfg3.jsonl 2>> $O/r2i_err.log
examples/_build/c_closed_loop 0 65536 300 1 2 4 8 > $O/r2i_c_closed_loop.json 2>> $O/r2i_err.log
CMD="python tools/bench_config.py --config 3 --steps 12 --warmup 4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lle_tiny_step_kernel -s 8 -c 1 -f -o $O/prof_cfg3_tiny6_r02 $CMD > $O/r2i_ncu_cfg3.log 2>&1
