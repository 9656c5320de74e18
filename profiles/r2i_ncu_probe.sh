mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 5 python -c "
import os
print({k:v for k,v in os.environ.items() if any(s in k.upper() for s in ('NV','CUDA','NSIGHT','INJECT','PRELOAD'))})
" 2>&1 | tail -5
for cfg in "default:" "oneStream:MOF_SMOOTH_AHEAD=0" "replay:MOF_MG_WHILE=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2i_$name.csv python tests/diag_timing.py 7 2 > gpurun_out/r2i_$name.log 2>&1; echo "$name rc $?"; tail -2 gpurun_out/r2i_$name.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2i_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/r2i_smoke.log
