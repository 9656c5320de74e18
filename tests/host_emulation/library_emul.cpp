// TEST INFRASTRUCTURE (CPU tier): the WHOLE library behind include/mof_b200.h compiled for the host — every .cu file of
// meshopticalflow_b200/csrc (dist.cu either stubbed as "one GPU", dist_stub.cpp, or over the in-process NCCL stand-in
// nccl.h, unit 7), kernels and host drivers alike, C ABI included —
// through emul_cuda_runtime.h: thread blocks on fibers, counted barriers, warp shuffles, atomics, stream capture and
// graph replay; a cooperative kernel runs as one CTA. One translation unit per source file, like the GPU build.
// Built by tests/test_library_host_emulation.py as: g++ -DMOF_HOST_EMULATION -DEMUL_UNIT=<n> library_emul.cpp ...
#include "emul_cuda_runtime.h"

#if EMUL_UNIT == 0
#include "../../meshopticalflow_b200/csrc/setup_kernels.cu"
#elif EMUL_UNIT == 1
#include "../../meshopticalflow_b200/csrc/pcg_kernels.cu"
#elif EMUL_UNIT == 2
#include "../../meshopticalflow_b200/csrc/multigrid.cu"
#elif EMUL_UNIT == 3
#include "../../meshopticalflow_b200/csrc/flow_kernels.cu"
#elif EMUL_UNIT == 4
#include "../../meshopticalflow_b200/csrc/vector_fields.cu"
#elif EMUL_UNIT == 5
#include "../../meshopticalflow_b200/csrc/texprep_kernels.cu"
#elif EMUL_UNIT == 6
#include "../../meshopticalflow_b200/csrc/mof_api.cu"
#elif EMUL_UNIT == 8
#include "../../meshopticalflow_b200/csrc/reorder.cu"
#elif EMUL_UNIT == 9
#include "../../meshopticalflow_b200/csrc/spectrum.cu"
#elif EMUL_UNIT == 7  // instead of dist_stub.cpp, with -DMOF_EMUL_THREADS: NCCL as threads of this process (nccl.h here)
#include "../../meshopticalflow_b200/csrc/dist.cu"
#endif
