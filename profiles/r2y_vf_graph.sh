# Conformal / Connection PCG with its batches of iterations as a replayed graph (MOF_VF_GRAPH) against eager launches; the three captures r2x missed.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; tail -4 gpurun_out/r2y_pytest.log
for cfg in "graph:" "eager:MOF_VF_GRAPH=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  for mode in "1 0" "2 0"; do
    tag=$(echo $mode | tr ' ' '_')
    ( env $envs timeout 300 python tests/diag_timing.py 7 3 $mode ) > gpurun_out/r2y_l7_${name}_$tag.log 2>&1
    echo "== $name vfMode/cMode $mode"; grep -E "^it[0-9]" gpurun_out/r2y_l7_${name}_$tag.log | cut -c1-150
  done
done
( timeout 300 python tests/diag_timing.py 9 2 2 0 ) > gpurun_out/r2y_l9_graph_2_0.log 2>&1; grep -E "^it[0-9]" gpurun_out/r2y_l9_graph_2_0.log | cut -c1-150
export MOF_MG_WHILE=0 MOF_SMOOTH_AHEAD=0
CMD="python tests/diag_timing.py 9 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_coarse_apply -s 60 -c 12 -f -o gpurun_out/r2y_coarse $CMD > gpurun_out/r2y_ncu_coarse.log 2>&1; echo "coarse rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_restrict -s 40 -c 4 -f -o gpurun_out/r2y_residual_restrict $CMD > gpurun_out/r2y_ncu_rr.log 2>&1; echo "rr rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_sample -s 1 -c 1 -f -o gpurun_out/r2y_walk $CMD > gpurun_out/r2y_ncu_walk.log 2>&1; echo "walk rc $?"
ls gpurun_out/r2y*.ncu-rep
