set -x
mkdir -p gpurun_out
( MOF_MG_VERBOSE=1 MOF_MG_TAIL_TRACE=1 MOF_SMOOTH_AHEAD=0 timeout 300 python tests/diag_timing.py 9 2 ) > gpurun_out/r2d_trace.log 2>&1; echo "rc $?"
grep "mg tail\|small levels" gpurun_out/r2d_trace.log | tail -4
for cfg in "c1536_1s:MOF_SMOOTH_AHEAD=0" "c5000_1s:MOF_SMOOTH_AHEAD=0 MOF_MG_TAIL_CELLS=5000" "c0_1s:MOF_SMOOTH_AHEAD=0 MOF_MG_TAIL_CELLS=0" "c1536:" "c5000:MOF_MG_TAIL_CELLS=5000" "c0:MOF_MG_TAIL_CELLS=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r2d_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2d_l9_$name.log
  grep -E "^it[0-9]|rc " gpurun_out/r2d_l9_$name.log | tail -3
done
