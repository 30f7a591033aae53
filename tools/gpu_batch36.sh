#!/bin/bash
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r3i_bench20.json 2> $O/r3i_err.log
python bench.py > $O/r3i_bench_default.json 2>> $O/r3i_err.log
