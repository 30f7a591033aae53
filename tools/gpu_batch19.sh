#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2r_pytest.log
python tools/step_trace.py --config 2 >> $O/r2r_grid.jsonl 2>> $O/r2r_err.log
python tools/timeline.py --isolated > $O/r2r_timeline.jsonl 2>> $O/r2r_err.log
python tools/timeline.py --isolated --envs 16384 >> $O/r2r_timeline.jsonl 2>> $O/r2r_err.log
python tools/timeline.py --isolated --envs 2048 >> $O/r2r_timeline.jsonl 2>> $O/r2r_err.log
examples/_build/c_closed_loop 0 65536 300 1 4 8 > $O/r2r_cl.json 2>> $O/r2r_err.log
python tools/bench_config.py --config 3 --repeat 2 >> $O/r2r_cfg.jsonl 2>> $O/r2r_err.log
python tools/bench_config.py --config 1 --repeat 2 >> $O/r2r_cfg.jsonl 2>> $O/r2r_err.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs > $O/r2r_bench20.log 2>> $O/r2r_err.log
