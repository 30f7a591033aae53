#!/bin/bash
O=gpurun_out
C=examples/_build/c_closed_loop
export LLE_B200_NBUF=2
python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x > $O/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2n_pytest.log
python tools/step_trace.py --config 2 >> $O/r2n_grid.jsonl 2>> $O/r2n_err.log
LLE_B200_STEP_CTAS_PER_SM=1 python tools/step_trace.py --config 2 >> $O/r2n_grid.jsonl 2>> $O/r2n_err.log
LLE_B200_STEP_CTAS_PER_SM=3 python tools/step_trace.py --config 2 >> $O/r2n_grid.jsonl 2>> $O/r2n_err.log
$C 0 65536 300 1 2 4 8 > $O/r2n_cl.json 2>> $O/r2n_err.log
python tools/e2e_probe.py --parts 1 4 > $O/r2n_probe.json 2>> $O/r2n_err.log
python tools/bench_config.py --config 1 --repeat 2 >> $O/r2n_cfg.jsonl 2>> $O/r2n_err.log
python tools/bench_config.py --config 4 --repeat 1 >> $O/r2n_cfg.jsonl 2>> $O/r2n_err.log
