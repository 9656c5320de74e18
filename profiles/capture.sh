#!/usr/bin/env bash
# Profiling recipe of this repo (run under gpurun on one B200; see /opt/skills/guides/B200_PROFILING.md).
#   1. plain run of the exact command (must exit 0),
#   2. launch list of the same command: kernel launches with their device time (cold-cache, serialised) from the
#      steady state of a timed alignment (the first ~60 000 launches are warm-up steps),
#   3. `--set full` captures of the dominant kernels: the fp64 SpMV+dot of the flow PCG, the fp32 fine sweep of the
#      multigrid cycle, the level-1 stencil kernel and the walk kernel (pass "quick" as 2nd argument for the first only).
# Outputs land in gpurun_out/; summaries are copied into profiles/ (see profiles/README.md).
set -uo pipefail
TAG="${1:-r1}"
MODE="${2:-full}"
CMD="python bench.py --steps 1 --warmup 3"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 60000 -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_spmv_dot -s 200 -c 1 -o gpurun_out/spmv_${TAG} $CMD > gpurun_out/ncu_spmv_${TAG}.log 2>&1
echo "spmv capture rc=$?"
if [ "$MODE" = "full" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_fine_apply_flow -s 400 -c 1 -o gpurun_out/fine_${TAG} $CMD > gpurun_out/ncu_fine_${TAG}.log 2>&1
  echo "fine sweep capture rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:k_coarse_apply -s 2400 -c 10 -o gpurun_out/coarse_${TAG} $CMD > gpurun_out/ncu_coarse_${TAG}.log 2>&1
  echo "coarse stencil capture rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:k_walk_sample -s 2 -c 1 -o gpurun_out/walk_${TAG} $CMD > gpurun_out/ncu_walk_${TAG}.log 2>&1
  echo "walk capture rc=$?"
fi
ls -la gpurun_out | tail -14
