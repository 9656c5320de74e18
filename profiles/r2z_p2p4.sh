# Peer-memory halo exchange + all-gather (one fused kernel per exchange) against NCCL on 4 B200: 16.8M vertices, 3 iterations.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for cfg in "p2p:MOF_MG_VERBOSE=1 MOF_DIST_P2P=1" "nccl:MOF_DIST_P2P=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 tests/dist_worker.py 11 3 > gpurun_out/r2z_l11_4gpu_$name.log 2>&1; echo "$name rc $?"
  grep "\[dist\]" gpurun_out/r2z_l11_4gpu_$name.log | head -1; grep "^{" gpurun_out/r2z_l11_4gpu_$name.log | tail -1
done
