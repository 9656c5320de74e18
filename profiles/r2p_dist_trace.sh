mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 tests/diag_dist_trace.py 11 2 > gpurun_out/r2p_trace8.log 2>&1; echo "rc $?"
grep -E "^{|dist trace" gpurun_out/r2p_trace8.log
