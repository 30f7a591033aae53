#!/bin/bash
# final validation of the round-2 tree: smoke, the whole GPU suite, the driver's bench line, config 3, the compiled closed loop
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2y_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2y_smoke.log
python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2y_pytest.log
python bench.py --steps 20 --warmup 5 > $O/r2y_bench.json 2> $O/r2y_bench.err
python tools/bench_config.py --config 3 --repeat 2 > $O/r2y_cfg3.jsonl 2>> $O/r2y_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2y_ref.json 2>> $O/r2y_bench.err
