# Renumbering of badly numbered meshes (reorder.cu); PDL on/off on the same box; K contexts on one GPU.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_scale.py -m gpu -x -q -k "badly or midsize" > gpurun_out/r2u_pytest.log 2>&1; tail -4 gpurun_out/r2u_pytest.log
for cfg in "random_auto:MOF_SYNTH_NUMBERING=random" "random_off:MOF_SYNTH_NUMBERING=random MOF_REORDER=0" "subdivision_auto:MOF_SYNTH_NUMBERING=subdivision" "morton:"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env MOF_SMOOTH_AHEAD=0 MOF_VERBOSE_SETUP=0 $envs timeout 600 python tests/diag_timing.py 9 3 ) > gpurun_out/r2u_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2u_l9_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc |set_mesh|SpMV|ERROR" gpurun_out/r2u_l9_$name.log | tail -6 | cut -c1-160
done
for cfg in "pdl:" "plain:MOF_PDL=0" "pdl2:" "plain2:MOF_PDL=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2u_bench_$name.json 2> gpurun_out/r2u_bench_$name.err; echo "bench $name rc $?"; cut -c1-160 gpurun_out/r2u_bench_$name.json
done
timeout 600 python tests/diag_concurrent.py 9 2 3 > gpurun_out/r2u_concurrent2.txt 2>&1; cat gpurun_out/r2u_concurrent2.txt
timeout 600 python tests/diag_concurrent.py 9 3 3 > gpurun_out/r2u_concurrent3.txt 2>&1; cat gpurun_out/r2u_concurrent3.txt
