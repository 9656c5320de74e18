mkdir -p gpurun_out
for cfg in "fused:" "unfused:MOF_MG_FUSE_RESTRICT=0" "skip:MOF_MG_SKIP_CELLS=6144 MOF_MG_SKIP_CELLS_SCALAR=6144" "g1s12:MOF_MG_GAMMA_LEVELS=1 MOF_MG_COARSE_SWEEPS=1,2" "g1s122:MOF_MG_GAMMA_LEVELS=1 MOF_MG_COARSE_SWEEPS=1,2,2" "g1s22:MOF_MG_GAMMA_LEVELS=1 MOF_MG_COARSE_SWEEPS=2,2" "g0s22:MOF_MG_GAMMA_LEVELS=0 MOF_MG_COARSE_SWEEPS=2,2,2" "g2s112:MOF_MG_COARSE_SWEEPS=1,1,2,2"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env MOF_SMOOTH_AHEAD=0 $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r2l_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2l_l9_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc " gpurun_out/r2l_l9_$name.log | tail -2 | cut -c1-120
done
