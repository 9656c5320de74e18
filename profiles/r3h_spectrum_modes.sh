# mof_spectrum for the other two bases (inverse-diagonal preconditioner): iterations and time at 16 386 and 65 538 vertices, 6 and 20 pairs.
mkdir -p gpurun_out
( timeout 200 python tests/diag_spectrum.py 6 6 1 0 ) > gpurun_out/r3h_spectrum_modes.txt 2>&1
( timeout 200 python tests/diag_spectrum.py 6 20 2 0 ) >> gpurun_out/r3h_spectrum_modes.txt 2>&1
( timeout 200 python tests/diag_spectrum.py 7 6 1 0 ) >> gpurun_out/r3h_spectrum_modes.txt 2>&1
( timeout 200 python tests/diag_spectrum.py 7 20 2 2 ) >> gpurun_out/r3h_spectrum_modes.txt 2>&1
cat gpurun_out/r3h_spectrum_modes.txt | cut -c1-260
