set -x
mkdir -p gpurun_out
( MOF_MG_TAIL_TRACE=1 MOF_SMOOTH_AHEAD=0 timeout 300 python tests/diag_timing.py 9 2 ) > gpurun_out/r2c_trace.log 2>&1; echo "rc $?"
grep "mg tail" gpurun_out/r2c_trace.log | tail -4
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py -m gpu -x -q --durations=10 > gpurun_out/r2c_pytest_scale.log 2>&1; tail -25 gpurun_out/r2c_pytest_scale.log
