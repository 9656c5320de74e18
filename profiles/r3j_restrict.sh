# k_restrict_flow with the edge vectors in member-list order (streamed) instead of gathered.
mkdir -p gpurun_out
timeout 300 python tests/diag_kernels.py 9 > gpurun_out/r3j_kernels.txt 2>&1; grep "flow_" gpurun_out/r3j_kernels.txt
( MOF_SMOOTH_AHEAD=0 timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r3j_l9_1s.log 2>&1; grep -E "^it[0-9]" gpurun_out/r3j_l9_1s.log | tail -3 | cut -c1-140
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_scale.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --quick > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; echo "bench rc $?"; cut -c1-170 gpurun_out/r3j_bench.json
