# Final state of the round: the whole GPU suite, smoke, the bench line.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2_final_pytest.log 2>&1; tail -12 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2_final_bench.err; cut -c1-220 gpurun_out/r2_final_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "ref rc $?"; cut -c1-300 gpurun_out/r2_final_ref.json
