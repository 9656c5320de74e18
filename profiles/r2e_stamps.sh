mkdir -p gpurun_out
( MOF_MG_TAIL_TRACE=1 MOF_SMOOTH_AHEAD=0 timeout 300 python tests/diag_timing.py 9 1 ) > gpurun_out/r2e_trace.log 2>&1; echo "rc $?"
grep -A8 "mg tail flow" gpurun_out/r2e_trace.log | tail -9
grep -A8 "mg tail scalar" gpurun_out/r2e_trace.log | tail -9
