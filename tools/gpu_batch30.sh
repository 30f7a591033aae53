#!/bin/bash
O=gpurun_out
python -c "import __graft_entry__ as g; g.build_c_client('c_closed_loop')" >> $O/r3c_err.log 2>&1
echo copy >> $O/r3c_loop.json
timeout 120 examples/_build/c_closed_loop 0 65536 300 s4 s6 s8 s12 >> $O/r3c_loop.json 2>> $O/r3c_err.log; echo "rc=$?" >> $O/r3c_err.log
echo zerocopy >> $O/r3c_loop.json
LLE_B200_PARTS_ZEROCOPY=1 timeout 120 examples/_build/c_closed_loop 0 65536 300 s4 s6 s8 s12 s16 >> $O/r3c_loop.json 2>> $O/r3c_err.log; echo "rc=$?" >> $O/r3c_err.log
