#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest.log
for e in 4 8 16; do
  LLE_B200_TINY_E=$e python tools/bench_config.py --config 3 --repeat 2 >> $O/r2d_cfg3.jsonl 2>> $O/r2d_err.log
done
python tools/step_trace.py --config 2 >> $O/r2d_grid.jsonl 2>> $O/r2d_err.log
LLE_B200_NARROW_DEPTH=2 python tools/step_trace.py --config 2 >> $O/r2d_grid.jsonl 2>> $O/r2d_err.log
python tools/e2e_probe.py > $O/r2d_e2e_probe.json 2>> $O/r2d_err.log
CMD="python tools/bench_config.py --config 3 --steps 12 --warmup 4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lle_tiny_step_kernel -s 8 -c 1 -f -o $O/prof_cfg3_tiny2_r02 $CMD > $O/r2d_ncu_cfg3.log 2>&1
