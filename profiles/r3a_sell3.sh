# Scalar fine sweep: three lanes per row (k_fine_apply_scalar_sell3) against one lane per row (MOF_SCALAR_SELL3=0).
mkdir -p gpurun_out
for cfg in "sell3:" "sell1:MOF_SCALAR_SELL3=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tests/diag_kernels.py 9 > gpurun_out/r3a_kernels_$name.txt 2>&1; echo "== $name"; grep scalar gpurun_out/r3a_kernels_$name.txt
  ( env $envs MOF_SMOOTH_AHEAD=0 timeout 300 python tests/diag_timing.py 9 10 ) > gpurun_out/r3a_l9_${name}_1s.log 2>&1; grep -E "^it[0-9]" gpurun_out/r3a_l9_${name}_1s.log | tail -2 | cut -c1-140
  env $envs timeout 600 python bench.py --steps 4 --warmup 3 --quick > gpurun_out/r3a_bench_$name.json 2> gpurun_out/r3a_bench_$name.err; echo "bench rc $?"; cut -c1-150 gpurun_out/r3a_bench_$name.json
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_scale.py -m gpu -x -q 2>&1 | tail -3
timeout 600 env MOF_MG_WHILE=0 MOF_SMOOTH_AHEAD=0 ncu --set full --clock-control none --import-source on -k regex:k_fine_apply_scalar_sell3 -s 40 -c 2 -f -o gpurun_out/r3a_scalar_sell3 python tests/diag_timing.py 9 1 > gpurun_out/r3a_ncu.log 2>&1; echo "ncu rc $?"
