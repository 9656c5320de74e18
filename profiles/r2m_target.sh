mkdir -p gpurun_out
for cfg in "t40:" "t80:MOF_MG_TARGET=80" "t160:MOF_MG_TARGET=160" "t320:MOF_MG_TARGET=320" "t160s56:MOF_MG_TARGET=160 MOF_MG_TARGET_SCALAR=56"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env MOF_SMOOTH_AHEAD=0 MOF_MG_VERBOSE=1 $envs timeout 300 python tests/diag_timing.py 9 4 ) > gpurun_out/r2m_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2m_l9_$name.log
  echo "== $name"; grep "mg flow\] level" gpurun_out/r2m_l9_$name.log | head -3; grep -E "^it[0-9]|rc " gpurun_out/r2m_l9_$name.log | tail -3 | cut -c1-120
done
