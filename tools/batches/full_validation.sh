#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider > $O/final_pytest.log 2>&1; echo "pytest rc=$?" >> $O/final_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/final_bench20b.json 2>> $O/final_err.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1
