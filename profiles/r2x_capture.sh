# --set full captures of the round-2 build's large kernels (one B200). The PCG runs as the replay loop (MOF_MG_WHILE=0: kernels inside the body of a
# conditional graph node are not profiled one by one) on one stream; the command is tests/diag_timing.py 9 1 (one UpdateFlow at 1 048 578 vertices),
# which exits 0 without ncu first.
mkdir -p gpurun_out
export MOF_MG_WHILE=0 MOF_SMOOTH_AHEAD=0
CMD="python tests/diag_timing.py 9 1"
timeout 300 $CMD > gpurun_out/r2x_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2x_plain.log; exit 1; }
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/r2x_$1 $CMD > gpurun_out/r2x_ncu_$1.log 2>&1; echo "$1 rc $?"
}
cap spmv k_spmv_dot 20 1
cap fine_flow "k_fine_apply_flow" 40 2
cap coarse_flow "k_coarse_apply<9" 60 12
cap residual_restrict "k_residual_restrict<9" 40 4
cap restrict_flow k_restrict_flow 20 1
cap update k_update_xr 20 2
cap scalar_sell k_fine_apply_scalar_sell 40 2
cap walk k_walk_sample 2 1
ls -la gpurun_out/*.ncu-rep | head -20
