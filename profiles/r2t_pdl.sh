# Programmatic dependent launches in the solvers (MOF_PDL, default on) against plain stream order; unsorted input numbering.
mkdir -p gpurun_out
for cfg in "pdl_1s:MOF_SMOOTH_AHEAD=0" "plain_1s:MOF_SMOOTH_AHEAD=0 MOF_PDL=0" "pdl_2s:" "plain_2s:MOF_PDL=0" "pdl_replay_1s:MOF_SMOOTH_AHEAD=0 MOF_MG_WHILE=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r2t_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2t_l9_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc |ERROR|rror" gpurun_out/r2t_l9_$name.log | tail -3 | cut -c1-140
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_scale.py tests/test_gpu_modes.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; tail -4 gpurun_out/r2t_pytest.log
for cfg in "subdivision:MOF_SYNTH_NUMBERING=subdivision" "random:MOF_SYNTH_NUMBERING=random"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env MOF_SMOOTH_AHEAD=0 $envs timeout 600 python tests/diag_timing.py 9 4 ) > gpurun_out/r2t_l9_numbering_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2t_l9_numbering_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc |set_mesh|SpMV|ERROR" gpurun_out/r2t_l9_numbering_$name.log | tail -6 | cut -c1-160
done
timeout 600 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc $?"; cut -c1-200 gpurun_out/r2t_bench.json
