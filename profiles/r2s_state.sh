# State of the build after the container was re-created: GPU suite, bench line, per-iteration times, kernel table, launch lists.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2s_pytest.log 2>&1; tail -14 gpurun_out/r2s_pytest.log
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc $?"; tail -c 400 gpurun_out/r2s_bench.err; cut -c1-600 gpurun_out/r2s_bench.json
for cfg in "1s:MOF_SMOOTH_AHEAD=0" "2s:"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r2s_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2s_l9_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc " gpurun_out/r2s_l9_$name.log | tail -3 | cut -c1-140
done
timeout 300 python tests/diag_kernels.py 9 > gpurun_out/r2s_kernels.txt 2>&1; cat gpurun_out/r2s_kernels.txt
for cfg in "while:" "replay:MOF_MG_WHILE=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 9000 --csv --log-file gpurun_out/r2s_launches_$name.csv python tests/diag_timing.py 9 2 > gpurun_out/r2s_ncu_$name.log 2>&1; echo "ncu $name rc $?"; tail -2 gpurun_out/r2s_ncu_$name.log | cut -c1-200
  python profiles/by_grid.py gpurun_out/r2s_launches_$name.csv 5 > gpurun_out/r2s_launches_${name}_by_grid.txt 2>&1; head -12 gpurun_out/r2s_launches_${name}_by_grid.txt
done
