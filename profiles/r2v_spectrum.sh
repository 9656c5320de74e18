# The Spectrum tool on the GPU: tests, then iterations / time at 65k and 1M vertices.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_spectrum.py -m gpu -x -q --durations=5 > gpurun_out/r2v_pytest.log 2>&1; tail -12 gpurun_out/r2v_pytest.log
( timeout 600 python tests/diag_spectrum.py 7 20 ) > gpurun_out/r2v_spectrum.txt 2>&1
( MOF_SPECTRUM_MG=1 timeout 300 python tests/diag_spectrum.py 9 20 2>&1 | head -1 ) >> gpurun_out/r2v_spectrum.txt 2>&1
( timeout 300 python tests/diag_spectrum.py 7 20 2 0 ) >> gpurun_out/r2v_spectrum.txt 2>&1
cat gpurun_out/r2v_spectrum.txt
