mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2h_pytest.log 2>&1; tail -16 gpurun_out/r2h_pytest.log
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc $?"; tail -c 600 gpurun_out/r2h_bench.err
timeout 600 python tests/diag_texprep.py > gpurun_out/r2h_texprep.txt 2>&1; tail -12 gpurun_out/r2h_texprep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fine_apply_scalar_row -s 60 -c 2 -o gpurun_out/r2h_scalar_row_before python tests/diag_timing.py 9 1 > gpurun_out/r2h_ncu_scalar.log 2>&1; echo "ncu rc $?"
