#!/bin/bash
# Runs on the GPU box (gpurun -- bash profiles/capture.sh TAG): the plain bench first (its numbers are the only ones reported),
# then the ncu launch list and full captures of the step kernel for the same workload.  Outputs land in gpurun_out/.
TAG=${1:-r02}
CMD="python bench.py --steps 64 --warmup 8 --no-cpu-baseline --e2e-steps 8 --no-configs --no-compiled-host"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "short bench failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
# three single-step launches of the headline kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lle_world_kernel -s 40 -c 3 -f -o gpurun_out/prof_${TAG}_final $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
# one lle_vec_rollout(128) launch: 128 steps in one kernel (duration / 128 and dram bytes / 128 corroborate the per-step figures)
RCMD="python tools/rollout_once.py"
timeout 900 ncu --set full --clock-control none -k regex:lle_world_kernel -s 6 -c 1 -f -o gpurun_out/prof_${TAG}_rollout128 $RCMD > gpurun_out/ncu_rollout_${TAG}.log 2>&1
echo "rollout capture rc=$?"
