#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "parts or pipelined" > $O/r3h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r3h_pytest.log
python -c "import __graft_entry__ as g; g.build_c_client('c_closed_loop')" >> $O/r3h_err.log 2>&1
timeout 120 examples/_build/c_closed_loop 0 65536 300 s8 s4 s16 8 > $O/r3h_loop.json 2>> $O/r3h_err.log; echo "rc=$?" >> $O/r3h_err.log
python tools/bench_config.py --config 3 --repeat 2 > $O/r3h_cfg.jsonl 2>> $O/r3h_err.log
python tools/bench_config.py --config 2 --repeat 2 >> $O/r3h_cfg.jsonl 2>> $O/r3h_err.log
