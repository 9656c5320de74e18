mkdir -p gpurun_out
( MOF_MG_VERBOSE=1 MOF_MG_TAIL_TRACE=1 MOF_SMOOTH_AHEAD=0 timeout 300 python tests/diag_timing.py 9 1 ) > gpurun_out/r2f_trace.log 2>&1; echo "rc $?"
grep "small levels" gpurun_out/r2f_trace.log | tail -2
grep "mg tail" gpurun_out/r2f_trace.log | tail -2
for cfg in "t6144_1s:MOF_SMOOTH_AHEAD=0" "t1536_1s:MOF_SMOOTH_AHEAD=0 MOF_MG_TAIL_CELLS=1536" "t0_1s:MOF_SMOOTH_AHEAD=0 MOF_MG_TAIL_CELLS=0" "t6144:" "t0:MOF_MG_TAIL_CELLS=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r2f_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2f_l9_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc " gpurun_out/r2f_l9_$name.log | tail -3 | cut -c1-120
done
