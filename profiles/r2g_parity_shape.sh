mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py -m gpu -q --durations=10 > gpurun_out/r2g_pytest_scale.log 2>&1; tail -22 gpurun_out/r2g_pytest_scale.log
for cfg in "g1:MOF_MG_GAMMA_LEVELS=1" "g3:MOF_MG_GAMMA_LEVELS=3" "fs2:MOF_MG_FINE_SWEEPS=2" "g1fs2:MOF_MG_GAMMA_LEVELS=1 MOF_MG_FINE_SWEEPS=2" "sfs2:MOF_MG_FINE_SWEEPS_SCALAR=2" "sg0:MOF_MG_GAMMA_LEVELS_SCALAR=0" "sg2:MOF_MG_GAMMA_LEVELS_SCALAR=2"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env MOF_SMOOTH_AHEAD=0 $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r2g_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2g_l9_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc " gpurun_out/r2g_l9_$name.log | tail -3 | cut -c1-120
done
