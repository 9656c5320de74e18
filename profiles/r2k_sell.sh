mkdir -p gpurun_out
python tests/diag_kernels.py 9 > gpurun_out/r2k_kernels_sell.txt 2>&1; cat gpurun_out/r2k_kernels_sell.txt
MOF_SCALAR_SELL=0 python tests/diag_kernels.py 9 > gpurun_out/r2k_kernels_csr.txt 2>&1; grep scalar gpurun_out/r2k_kernels_csr.txt
for cfg in "sell_1s:MOF_SMOOTH_AHEAD=0" "csr_1s:MOF_SMOOTH_AHEAD=0 MOF_SCALAR_SELL=0" "sell:" ; do
  name=${cfg%%:*}; envs=${cfg#*:}
  ( env $envs timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r2k_l9_$name.log 2>&1; echo "rc $?" >> gpurun_out/r2k_l9_$name.log
  echo "== $name"; grep -E "^it[0-9]|rc " gpurun_out/r2k_l9_$name.log | tail -2 | cut -c1-120
done
timeout 600 env MOF_MG_WHILE=0 ncu --set full --clock-control none --import-source on -k regex:k_fine_apply_scalar_sell -s 40 -c 2 -o gpurun_out/r2k_scalar_sell python tests/diag_timing.py 9 1 > gpurun_out/r2k_ncu_sell.log 2>&1; echo "ncu sell rc $?"
timeout 600 env MOF_MG_WHILE=0 MOF_SCALAR_SELL=0 ncu --set full --clock-control none --import-source on -k regex:k_fine_apply_scalar_row -s 40 -c 2 -o gpurun_out/r2k_scalar_row python tests/diag_timing.py 9 1 > gpurun_out/r2k_ncu_row.log 2>&1; echo "ncu row rc $?"
