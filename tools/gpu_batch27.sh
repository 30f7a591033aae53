#!/bin/bash
# partial observations on the thread-per-world kernel: parity, then level 6 x 65,536 against the general kernel
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -x -k "observation or partial or config2 or config3" > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2z_pytest.log
for ot in partial3x3 partial5x5 partial7x7; do
  echo "tiny $ot" >> $O/r2z_partial.jsonl
  python tools/bench_config.py --config 2 --obs-type $ot --repeat 2 >> $O/r2z_partial.jsonl 2>> $O/r2z_err.log
  echo "general $ot" >> $O/r2z_partial.jsonl
  LLE_B200_NO_TINY=1 python tools/bench_config.py --config 2 --obs-type $ot --repeat 2 >> $O/r2z_partial.jsonl 2>> $O/r2z_err.log
done
for e in 2 8; do
  echo "tiny partial3x3 E=$e" >> $O/r2z_partial.jsonl
  LLE_B200_TINY_E=$e python tools/bench_config.py --config 2 --obs-type partial3x3 >> $O/r2z_partial.jsonl 2>> $O/r2z_err.log
done
for e in 1 4; do
  echo "tiny partial5x5 E=$e" >> $O/r2z_partial.jsonl
  LLE_B200_TINY_E=$e python tools/bench_config.py --config 2 --obs-type partial5x5 >> $O/r2z_partial.jsonl 2>> $O/r2z_err.log
done
python tools/bench_config.py --config 3 --repeat 2 >> $O/r2z_partial.jsonl 2>> $O/r2z_err.log
python bench.py --steps 20 --warmup 5 > $O/r2z_bench.json 2>> $O/r2z_err.log
