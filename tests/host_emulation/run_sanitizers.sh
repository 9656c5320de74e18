#!/bin/bash
# TEST INFRASTRUCTURE: memory and undefined-behaviour checks of every kernel without a GPU. The whole-library and the
# texture-preparation and the partitioned-mesh emulation tests are run twice:
#  1. AddressSanitizer — "device" memory of the emulated build is malloc'd, so a kernel that reads or writes past a buffer fails the
#     test that ran it, with the .cu file and line in the report (compute-sanitizer memcheck's job on a GPU box);
#  2. UndefinedBehaviorSanitizer — signed overflow in index arithmetic, bad shifts, and misaligned vector loads: float2 / double2 /
#     float4 carry the device's alignment in emul_cuda_runtime.h, so a load that would be a "misaligned address" fault on the GPU aborts.
#   tests/host_emulation/run_sanitizers.sh            (about 9 minutes)
set -e
cd "$(dirname "$0")/../.."
TESTS="tests/test_library_host_emulation.py tests/test_texprep_host_emulation.py tests/test_dist_host_emulation.py"
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
MOF_EMUL_CXXFLAGS="-O1 -g -fsanitize=address -fno-omit-frame-pointer" python -m pytest $TESTS -x -q "$@"
MOF_EMUL_CXXFLAGS="-O1 -g -fsanitize=undefined -fno-sanitize-recover=undefined" python -m pytest $TESTS -x -q "$@"
