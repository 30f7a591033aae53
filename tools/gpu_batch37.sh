#!/bin/bash
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r3j_bench20a.json 2> $O/r3j_err.log
python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r3j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r3j_pytest.log
python bench.py --steps 20 --warmup 5 > $O/r3j_bench20b.json 2>> $O/r3j_err.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r3j_smoke.log 2>&1
