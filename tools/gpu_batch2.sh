#!/bin/bash
# GPU batch 2 (round 2): full GPU suite, step-grid sweep, tiny-kernel A/B on config 3, the driver's bench command.
O=gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
for c in 1 2 3 4 5; do
  LLE_B200_FORCE_NARROW=1 LLE_B200_STEP_CTAS_PER_SM=$c python tools/step_trace.py --config 2 >> $O/r2b_grid_sweep.jsonl 2>> $O/r2b_err.log
done
python tools/step_trace.py --config 2 >> $O/r2b_grid_sweep.jsonl 2>> $O/r2b_err.log
LLE_B200_NO_TINY=1 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2b_cfg3.jsonl 2>> $O/r2b_err.log
for e in 4 8 16 32; do
  LLE_B200_TINY_E=$e python tools/bench_config.py --config 3 --repeat 2 >> $O/r2b_cfg3.jsonl 2>> $O/r2b_err.log
done
for c in 4 8 12; do
  LLE_B200_TINY_E=8 LLE_B200_TINY_CTAS_PER_SM=$c python tools/bench_config.py --config 3 --repeat 2 >> $O/r2b_cfg3.jsonl 2>> $O/r2b_err.log
done
python tools/bench_config.py --config 1 >> $O/r2b_cfg1.jsonl 2>> $O/r2b_err.log
( time python bench.py --steps 20 --warmup 5 ) > $O/r2b_bench20.log 2> $O/r2b_bench20.err
python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-configs > $O/r2b_bench2048.log 2>> $O/r2b_err.log
