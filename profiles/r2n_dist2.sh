mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
for cfg in "dealt:" "replicated:MOF_DIST_LEVEL_CELLS=100000000"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs MOF_MG_VERBOSE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_worker.py 10 3 > gpurun_out/r2n_l10_2gpu_$name.log 2>&1; echo "$name rc $?"
  grep "^{" gpurun_out/r2n_l10_2gpu_$name.log | tail -1
done
