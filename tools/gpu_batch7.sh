#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log
for e in 4 8 16 32; do
  LLE_B200_TINY_E=$e python tools/bench_config.py --config 3 --repeat 2 >> $O/r2g_cfg3.jsonl 2>> $O/r2g_err.log
done
python tools/bench_config.py --config 1 --repeat 2 >> $O/r2g_cfg3.jsonl 2>> $O/r2g_err.log
examples/_build/c_closed_loop 0 65536 300 1 2 4 8 > $O/r2g_c_closed_loop.json 2>> $O/r2g_err.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs > $O/r2g_bench20.log 2>> $O/r2g_err.log
CMD="python tools/bench_config.py --config 3 --steps 12 --warmup 4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lle_tiny_step_kernel -s 8 -c 1 -f -o $O/prof_cfg3_tiny5_r02 $CMD > $O/r2g_ncu_cfg3.log 2>&1
