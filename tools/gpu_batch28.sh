#!/bin/bash
# partial 3x3 on the thread-per-world kernel (default E), parity of the forced large windows; closed loop with capped grids
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -x -k "observation or partial or windows or config2" > $O/r3a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r3a_pytest.log
python tools/bench_config.py --config 2 --obs-type partial3x3 --repeat 2 >> $O/r3a_partial.jsonl 2>> $O/r3a_err.log
python -c "import __graft_entry__ as g; g.build_c_client('c_closed_loop')" >> $O/r3a_err.log 2>&1
for cap in 0 16 32 48 64 96; do
  echo "cap $cap" >> $O/r3a_loop.jsonl
  LLE_B200_GRID_CAP=$cap examples/_build/c_closed_loop 0 65536 300 4 8 16 >> $O/r3a_loop.jsonl 2>> $O/r3a_err.log
done
