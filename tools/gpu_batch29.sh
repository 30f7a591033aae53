#!/bin/bash
O=gpurun_out
python -c "import __graft_entry__ as g; g.build_c_client('c_closed_loop')" >> $O/r3b_err.log 2>&1
timeout 120 examples/_build/c_closed_loop 0 65536 300 s8 8 s4 s16 s32 > $O/r3b_loop.json 2>> $O/r3b_err.log; echo "rc=$?" >> $O/r3b_err.log
timeout 120 examples/_build/c_closed_loop 0 65536 300 s8 s16 >> $O/r3b_loop.json 2>> $O/r3b_err.log; echo "rc=$?" >> $O/r3b_err.log
python tools/bench_config.py --config 3 --repeat 2 > $O/r3b_cfg3.jsonl 2>> $O/r3b_err.log
python tools/bench_config.py --config 2 --repeat 2 >> $O/r3b_cfg3.jsonl 2>> $O/r3b_err.log
