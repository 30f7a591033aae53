#!/bin/bash
# Runs on the GPU box (gpurun -- bash profiles/capture.sh TAG): the plain bench first (its numbers are the only ones reported),
# then the ncu launch list and one full capture of the step kernel for the same workload.  Outputs land in gpurun_out/.
TAG=${1:-r01}
CMD="python bench.py --steps 64 --warmup 8 --no-cpu-baseline --e2e-steps 8"
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err || { echo "bench failed"; tail -5 gpurun_out/bench_${TAG}.err; exit 1; }
python bench.py --impl reference > gpurun_out/bench_ref_${TAG}.log 2>> gpurun_out/bench_${TAG}.err
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "short bench failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lle_world_kernel -s 40 -c 3 -f -o gpurun_out/prof_${TAG}_final $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
cut -c1-300 gpurun_out/bench_${TAG}.log
