#!/bin/bash
# long closed loops over parts at several part counts, launch depths and a ragged batch size: the host compares every done flag of
# every part and step with the recording (any lost / reordered / stale result would show) - examples/c_closed_loop.c checks itself
O=gpurun_out
python -c "import __graft_entry__ as g; g.build_c_client('c_closed_loop')" >> $O/stress_err.log 2>&1
for ahead in 2 3 4 1; do
  echo "ahead=$ahead n=65536" >> $O/stress.jsonl
  LLE_LOOP_AHEAD=$ahead timeout 300 examples/_build/c_closed_loop 0 65536 3000 s4 s16 s64 s1 >> $O/stress.jsonl 2>> $O/stress_err.log; echo "rc=$?" >> $O/stress.jsonl
done
echo "ahead=2 n=65003" >> $O/stress.jsonl
timeout 300 examples/_build/c_closed_loop 0 65003 3000 s7 s13 >> $O/stress.jsonl 2>> $O/stress_err.log; echo "rc=$?" >> $O/stress.jsonl
echo "ahead=2 n=1000 (one CTA wave, many parts)" >> $O/stress.jsonl
timeout 300 examples/_build/c_closed_loop 0 1000 20000 s5 s50 >> $O/stress.jsonl 2>> $O/stress_err.log; echo "rc=$?" >> $O/stress.jsonl
