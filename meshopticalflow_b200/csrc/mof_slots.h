// Constants shared by the host code and the kernels (no CUDA dependence).
#pragma once

namespace mof {

#ifdef MOF_HOST_EMULATION
constexpr int kSMs = 1;    // CPU test tier (tests/host_emulation): every "CTA" is run thread by thread
#else
constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
#endif

// Device-side scalar slots (one small buffer, fp64).
enum ScalarSlot {
    SC_AREA_SCALE = 0,   // 2 / sum sqrt det g
    SC_FROB2 = 1,        // ||R D P||_F^2
    SC_DATA_SCALE = 2,   // 1 / ||R D P||_F
    SC_STEP_NUM = 3,     // x . b
    SC_STEP_DEN = 4,     // x . Dt x
    SC_TMP = 5,          // scratch of the call in progress
    SC_DOG = 8,          // 8..8+4*6: per channel old avg, old dot, new avg, new dot
    SC_COUNT = 64
};

}  // namespace mof
