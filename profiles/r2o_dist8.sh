mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tests/dist_worker.py 11 3 ab > gpurun_out/r2o_l11_8gpu.log 2>&1; echo "rc $?"
grep "^{" gpurun_out/r2o_l11_8gpu.log | tail -1; tail -3 gpurun_out/r2o_l11_8gpu.log | cut -c1-300
