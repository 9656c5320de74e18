# Final state: the whole GPU suite; the launch list of the bench command itself (replay loop: conditional graph bodies are not profiled per kernel).
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r3g_pytest.log 2>&1; tail -10 gpurun_out/r3g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
MOF_MG_WHILE=0 timeout 600 python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r3g_plain.json 2> gpurun_out/r3g_plain.err; echo "plain rc $?"; cut -c1-160 gpurun_out/r3g_plain.json
MOF_MG_WHILE=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100000 -c 12000 --csv --log-file gpurun_out/r3g_launches.csv python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r3g_ncu.log 2>&1; echo "ncu rc $?"
python profiles/by_grid.py gpurun_out/r3g_launches.csv 5 > gpurun_out/r3g_launches_by_grid.txt 2>&1; head -16 gpurun_out/r3g_launches_by_grid.txt
python profiles/summarize.py launches gpurun_out/r3g_launches.csv > gpurun_out/r3g_launches.txt 2>&1; head -12 gpurun_out/r3g_launches.txt
