mkdir -p gpurun_out
for cfg in "p10:MOF_MG_POWER_ITS=10" "p5:MOF_MG_POWER_ITS=5" "p3:MOF_MG_POWER_ITS=3"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env MOF_SMOOTH_AHEAD=0 $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r2j_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2j_l9_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc " gpurun_out/r2j_l9_$name.log | tail -3 | cut -c1-120
done
