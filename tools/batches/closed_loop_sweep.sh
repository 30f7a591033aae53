#!/bin/bash
O=gpurun_out
python -c "import __graft_entry__ as g; g.build_c_client('c_closed_loop')" >> $O/r3g_err.log 2>&1
for ctas in 3 4 5; do for ahead in 2; do
  echo "ctas=$ctas ahead=$ahead" >> $O/r3g_loop.json
  LLE_B200_MAX_CTAS_PER_SM=$ctas LLE_LOOP_AHEAD=$ahead timeout 120 examples/_build/c_closed_loop 0 65536 300 s4 s8 s12 s16 s24 >> $O/r3g_loop.json 2>> $O/r3g_err.log; echo "rc=$?" >> $O/r3g_err.log
done; done
echo "ctas=3 ahead=1" >> $O/r3g_loop.json
LLE_B200_MAX_CTAS_PER_SM=3 LLE_LOOP_AHEAD=1 timeout 120 examples/_build/c_closed_loop 0 65536 300 s8 s16 >> $O/r3g_loop.json 2>> $O/r3g_err.log
