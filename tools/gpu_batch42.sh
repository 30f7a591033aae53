#!/bin/bash
O=gpurun_out
python tools/bench_config.py --config 3 --repeat 2 > $O/r3p_cfg3.jsonl 2>> $O/r3p_err.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider -k "config3 or levels or corpus or smoke or options" > $O/r3p_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r3p_pytest.log
python tools/bench_config.py --config 3 --repeat 2 >> $O/r3p_cfg3.jsonl 2>> $O/r3p_err.log
