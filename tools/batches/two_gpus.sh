#!/bin/bash
O=gpurun_out
nvidia-smi -L > $O/two_2gpu.log 2>&1
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -k "two_gpus" >> $O/two_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/two_2gpu.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/two_bench_2gpu.json 2> $O/two_bench_2gpu.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > $O/two_ref_2gpu.json 2>> $O/two_bench_2gpu.err
