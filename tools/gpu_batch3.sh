#!/bin/bash
# GPU batch 3: ncu of the tiny kernel on config 3; narrow-grid policy A/B at 20-step windows
O=gpurun_out
CMD="python tools/bench_config.py --config 3 --steps 12 --warmup 4"
$CMD > $O/r2c_cfg3_plain.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lle_tiny_step_kernel -s 8 -c 1 -f -o $O/prof_cfg3_tiny_r02 $CMD > $O/r2c_ncu_cfg3.log 2>&1
echo "ncu rc=$?" >> $O/r2c_ncu_cfg3.log
for c in 1 2; do
  LLE_B200_STEP_CTAS_PER_SM=$c python tools/step_trace.py --config 2 >> $O/r2c_grid.jsonl 2>> $O/r2c_err.log
done
