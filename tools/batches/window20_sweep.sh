#!/bin/bash
# the driver's 20-step window against the width of the launch that finds the device idle (LLE_B200_MAX_CTAS_PER_SM) and of the
# launches that find a predecessor in flight (LLE_B200_STEP_CTAS_PER_SM)
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-compiled-host --no-closed-loop --e2e-steps 4"
for rep in 1 2; do
for cfg in "5 2" "3 2" "4 2" "2 2" "3 3" "3 1"; do
  set -- $cfg
  echo "idle=$1 busy=$2" >> $O/r3q_win.jsonl
  LLE_B200_MAX_CTAS_PER_SM=$1 LLE_B200_STEP_CTAS_PER_SM=$2 $B 2>> $O/r3q_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'ms': d['ms_per_step'], 'sync': d['e2e']['sync_value']}))" >> $O/r3q_win.jsonl
done; done
