# Peer-memory halo exchange (MOF_DIST_P2P) against NCCL send/recv on 2 B200: parity test, then 4.2M vertices, 3 iterations.
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
for cfg in "p2p:MOF_MG_VERBOSE=1 MOF_DIST_P2P=1" "nccl:MOF_DIST_P2P=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tests/dist_worker.py 10 3 > gpurun_out/r2z_l10_2gpu_$name.log 2>&1; echo "$name rc $?"
  grep "\[dist\]" gpurun_out/r2z_l10_2gpu_$name.log | head -2; grep "^{" gpurun_out/r2z_l10_2gpu_$name.log | tail -1
done
