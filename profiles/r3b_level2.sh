# Level 2 of the flow hierarchy (18.9 k cells at 1M vertices): the 27-warp stencil variant and the fused residual+restriction, which today start below 16 384 cells.
mkdir -p gpurun_out
for cfg in "default:" "wide32k:MOF_MG_WIDE_FROM=32768" "fuse32k:MOF_MG_FUSE_BELOW=32768 MOF_MG_WIDE_FROM=32768" "wide128k:MOF_MG_WIDE_FROM=131072" "fuse128k:MOF_MG_FUSE_BELOW=131072 MOF_MG_WIDE_FROM=131072"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env MOF_SMOOTH_AHEAD=0 $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r3b_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r3b_l9_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc |rror" gpurun_out/r3b_l9_$name.log | tail -3 | cut -c1-140
done
