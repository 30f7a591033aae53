#!/bin/bash
O=gpurun_out
C=examples/_build/c_closed_loop
$C 0 65536 300 4 8 16 > $O/r2m_cl_default.json 2>> $O/r2m_err.log
for c in 1 2 3; do
  LLE_B200_FORCE_NARROW=1 LLE_B200_STEP_CTAS_PER_SM=$c $C 0 65536 300 2 4 8 > $O/r2m_cl_narrow$c.json 2>> $O/r2m_err.log
done
python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2m_pytest.log
python bench.py --steps 20 --warmup 5 > $O/r2m_bench20.log 2>> $O/r2m_err.log
