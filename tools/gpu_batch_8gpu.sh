#!/bin/bash
O=gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 ) > $O/r2q_bench_8gpu.json 2> $O/r2q_bench_8gpu.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 ) > $O/r2q_bench_4gpu.json 2>> $O/r2q_bench_8gpu.err
