#!/bin/bash
O=gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 ) > $O/r3t_bench_4gpu.json 2> $O/r3t_bench_4gpu.err
