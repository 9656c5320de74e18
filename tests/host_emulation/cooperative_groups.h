// TEST INFRASTRUCTURE (CPU tier): stand-in for <cooperative_groups.h>. A cooperative kernel is emulated with ONE CTA
// (emul_cuda_runtime.h), so the grid barrier is the CTA barrier.
#pragma once
#include "emul_cuda_runtime.h"

namespace cooperative_groups {
struct grid_group {
    void sync() const { __syncthreads(); }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
