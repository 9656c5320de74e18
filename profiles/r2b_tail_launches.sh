set -x
mkdir -p gpurun_out
export MOF_MG_WHILE=0 MOF_SMOOTH_AHEAD=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 4000 --csv --log-file gpurun_out/r2b_launches_tail.csv python tests/diag_timing.py 9 1 > gpurun_out/r2b_ncu_tail.log 2>&1
echo rc $?
python profiles/by_grid.py gpurun_out/r2b_launches_tail.csv 5 > gpurun_out/r2b_launches_tail_by_grid.txt
head -30 gpurun_out/r2b_launches_tail_by_grid.txt
