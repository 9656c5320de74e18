# mof_spectrum for the Conformal basis with the two-cycle preconditioner over the scalar hierarchy; the spectrum tests of the GPU tier.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spectrum.py -m gpu -x -q 2>&1 | tail -4
( timeout 200 python tests/diag_spectrum.py 6 6 1 0 ) > gpurun_out/r3i_spectrum_conformal.txt 2>&1
( timeout 300 python tests/diag_spectrum.py 7 20 1 0 ) >> gpurun_out/r3i_spectrum_conformal.txt 2>&1
( timeout 400 python tests/diag_spectrum.py 9 6 1 0 ) >> gpurun_out/r3i_spectrum_conformal.txt 2>&1
cat gpurun_out/r3i_spectrum_conformal.txt | cut -c1-260
