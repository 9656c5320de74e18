#!/usr/bin/env bash
# Profiling recipe of this repo (run under gpurun on one B200; see /opt/skills/guides/B200_PROFILING.md).
#   1. plain run of the exact command (must exit 0),
#   2. launch list of the same command: every kernel launch with its device time (cold-cache, serialised),
#   3. one `--set full` capture of the dominant kernel (k_spmv_dot = phase 1 of k_pcg<1>) and of the walk kernel.
# Outputs land in gpurun_out/; summaries are copied into profiles/ by hand (see profiles/README.md).
set -uo pipefail
TAG="${1:-r1}"
CMD="python bench.py --steps 1 --warmup 3"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_spmv_dot -s 10 -c 2 -o gpurun_out/spmv_${TAG} $CMD > gpurun_out/ncu_spmv_${TAG}.log 2>&1
echo "spmv capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_walk_sample -s 2 -c 1 -o gpurun_out/walk_${TAG} $CMD > gpurun_out/ncu_walk_${TAG}.log 2>&1
echo "walk capture rc=$?"
ls -la gpurun_out | tail -12
