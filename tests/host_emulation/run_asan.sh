#!/bin/bash
# TEST INFRASTRUCTURE: a memory check of every kernel without a GPU. "Device" memory of the emulated build is malloc'd, so with
# AddressSanitizer a kernel that reads or writes past a buffer fails the test that ran it, with the .cu file and line in the report
# (the role compute-sanitizer's memcheck plays on a GPU box). Runs the whole-library and the texture-preparation emulation tests.
#   tests/host_emulation/run_asan.sh            (about 2 minutes)
set -e
cd "$(dirname "$0")/../.."
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
MOF_EMUL_CXXFLAGS="-O1 -g -fsanitize=address -fno-omit-frame-pointer" \
python -m pytest tests/test_library_host_emulation.py tests/test_texprep_host_emulation.py -x -q "$@"
