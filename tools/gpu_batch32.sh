#!/bin/bash
O=gpurun_out
python -c "import __graft_entry__ as g; g.build_c_client('c_closed_loop')" >> $O/r3e_err.log 2>&1
for zc in 1 0; do for ctas in 2 1 3; do for ahead in 2 3; do
  echo "zerocopy=$zc ctas=$ctas ahead=$ahead" >> $O/r3e_loop.json
  LLE_B200_PARTS_ZEROCOPY=$zc LLE_B200_STEP_CTAS_PER_SM=$ctas LLE_LOOP_AHEAD=$ahead timeout 120 examples/_build/c_closed_loop 0 65536 300 s8 >> $O/r3e_loop.json 2>> $O/r3e_err.log; echo "rc=$?" >> $O/r3e_err.log
done; done; done
