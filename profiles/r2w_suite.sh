# Whole GPU suite on the build with renumbering + PDL + Spectrum; the bench line; Spectrum timings with the retuned preconditioner shift.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2w_pytest.log 2>&1; tail -14 gpurun_out/r2w_pytest.log
( MGS=1 timeout 300 python tests/diag_spectrum.py 7 20 ) > gpurun_out/r2w_spectrum.txt 2>&1
( MGS=1 timeout 600 python tests/diag_spectrum.py 9 20 ) >> gpurun_out/r2w_spectrum.txt 2>&1
cat gpurun_out/r2w_spectrum.txt
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc $?"; tail -c 300 gpurun_out/r2w_bench.err; cut -c1-200 gpurun_out/r2w_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print(d.get('concurrent_contexts')); print(d['e2e'])"
