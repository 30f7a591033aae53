#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -x > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2s_pytest.log
python tools/bench_config.py --config 3 --repeat 2 >> $O/r2s_cfg.jsonl 2>> $O/r2s_err.log
python tools/bench_config.py --config 2 --repeat 2 >> $O/r2s_cfg.jsonl 2>> $O/r2s_err.log
examples/_build/c_closed_loop 0 65536 300 1 4 8 > $O/r2s_cl.json 2>> $O/r2s_err.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2s_bench20.log 2>> $O/r2s_err.log
