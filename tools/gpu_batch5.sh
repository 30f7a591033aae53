#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log
for e in 4 8 16; do
  LLE_B200_TINY_E=$e python tools/bench_config.py --config 3 --repeat 2 >> $O/r2e_cfg3.jsonl 2>> $O/r2e_err.log
done
python tools/e2e_probe.py --parts 1 2 4 8 > $O/r2e_e2e_probe.json 2>> $O/r2e_err.log
python bench.py --steps 20 --warmup 5 > $O/r2e_bench20.log 2>> $O/r2e_err.log
CMD="python tools/bench_config.py --config 3 --steps 12 --warmup 4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lle_tiny_step_kernel -s 8 -c 1 -f -o $O/prof_cfg3_tiny3_r02 $CMD > $O/r2e_ncu_cfg3.log 2>&1
