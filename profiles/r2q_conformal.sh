mkdir -p gpurun_out
( MOF_MG_VERBOSE=1 timeout 300 python tests/diag_timing.py 7 2 1 ) 2>&1 | grep -E "conformal\]|^it[0-9]" | tail -6
( MOF_MG_VERBOSE=1 timeout 600 python tests/diag_timing.py 9 2 1 ) > gpurun_out/r2q_conformal_1M.txt 2>&1; echo "rc $?"; grep -E "conformal\]|^it[0-9]|set_signals|Error|error" gpurun_out/r2q_conformal_1M.txt | tail -8
timeout 600 python -m pytest tests/test_gpu_modes.py -m gpu -x -q 2>&1 | tail -3
