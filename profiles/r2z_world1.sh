# The partitioned code path with a world of ONE rank against the plain single-GPU path (same GPU, 4.2M vertices): what the path itself costs before any exchange.
mkdir -p gpurun_out
timeout 300 python tests/dist_worker.py 10 3 > gpurun_out/r2z_l10_world1.log 2>&1; echo "rc $?"; grep "^{" gpurun_out/r2z_l10_world1.log | tail -1
MOF_DIST_GRAPH=0 timeout 300 python tests/dist_worker.py 10 3 > gpurun_out/r2z_l10_world1_eager.log 2>&1; echo "rc $?"; grep "^{" gpurun_out/r2z_l10_world1_eager.log | tail -1
