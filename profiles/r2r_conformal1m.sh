mkdir -p gpurun_out
( MOF_MG_VERBOSE=1 MOF_VF_VERBOSE=1 timeout 400 python tests/diag_timing.py 9 3 1 ) > gpurun_out/r2r_conformal_1M.txt 2>&1; echo "rc $?"; grep -E "TRUE|conformal\]|^it|unknowns" gpurun_out/r2r_conformal_1M.txt | head -40
( timeout 300 python tests/diag_timing.py 9 2 2 ) 2>&1 | grep -E "^it|unknowns" | cut -c1-400
