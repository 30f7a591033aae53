#!/bin/bash
O=gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 ) > $O/r3m_bench_8gpu.json 2> $O/r3m_bench_8gpu.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/r3m_bench_2gpu.json 2>> $O/r3m_bench_8gpu.err
