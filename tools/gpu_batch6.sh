#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2f_pytest.log
for e in 4 8; do
  LLE_B200_TINY_E=$e python tools/bench_config.py --config 3 --repeat 2 >> $O/r2f_cfg3.jsonl 2>> $O/r2f_err.log
done
LLE_B200_TINY_E=4 LLE_B200_TINY_CTAS_PER_SM=4 python tools/bench_config.py --config 3 --repeat 2 >> $O/r2f_cfg3.jsonl 2>> $O/r2f_err.log
python tools/bench_config.py --config 1 --repeat 2 >> $O/r2f_cfg3.jsonl 2>> $O/r2f_err.log
python tools/bench_config.py --config 2 --repeat 2 >> $O/r2f_cfg3.jsonl 2>> $O/r2f_err.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2f_bench20.log 2>> $O/r2f_err.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --e2e-parts 8 > $O/r2f_bench20_p8.log 2>> $O/r2f_err.log
CMD="python tools/bench_config.py --config 3 --steps 12 --warmup 4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lle_tiny_step_kernel -s 8 -c 1 -f -o $O/prof_cfg3_tiny4_r02 $CMD > $O/r2f_ncu_cfg3.log 2>&1
