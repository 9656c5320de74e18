set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/r2a_info.txt
( timeout 120 python tests/diag_timing.py 7 3 ) > gpurun_out/r2a_l7.log 2>&1; echo "rc $?" >> gpurun_out/r2a_l7.log
tail -3 gpurun_out/r2a_l7.log
for cfg in "default:" "nowhile:MOF_MG_WHILE=0" "notail:MOF_MG_TAIL_CELLS=0" "old:MOF_MG_WHILE=0 MOF_MG_TAIL_CELLS=0" "tail0:MOF_MG_TAIL_CELLS=100000" "default_1s:MOF_SMOOTH_AHEAD=0" "old_1s:MOF_SMOOTH_AHEAD=0 MOF_MG_WHILE=0 MOF_MG_TAIL_CELLS=0" "tail0_1s:MOF_SMOOTH_AHEAD=0 MOF_MG_TAIL_CELLS=100000"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r2a_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2a_l9_$name.log
  grep -E "^it[0-9]|rc " gpurun_out/r2a_l9_$name.log | tail -4
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -5 gpurun_out/r2a_pytest.log
