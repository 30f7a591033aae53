#!/bin/bash
# Runs on the GPU box (gpurun -- bash profiles/capture_final.sh): the plain commands first (their numbers are the only ones reported),
# then the ncu launch list of the bench command and full captures of the headline kernel and of the thread-per-world kernel on
# config 3.  The closed loop over parts is skipped under ncu (bench.py: a serialising profiler would deadlock it).  -> gpurun_out/
TAG=r02f
CMD="python bench.py --steps 64 --warmup 8 --no-cpu-baseline --e2e-steps 8 --no-configs --no-compiled-host --no-closed-loop"
C3="python tools/bench_config.py --config 3 --steps 30 --warmup 6"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "short bench failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
$C3 > gpurun_out/plain_c3_${TAG}.log 2>&1 || { echo "config 3 failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lle_world_kernel -s 40 -c 3 -f -o gpurun_out/prof_${TAG}_final $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lle_tiny_step_kernel -s 12 -c 2 -f -o gpurun_out/prof_${TAG}_cfg3 $C3 > gpurun_out/ncu_cfg3_${TAG}.log 2>&1
echo "config 3 capture rc=$?"
