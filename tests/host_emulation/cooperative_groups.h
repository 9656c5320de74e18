// TEST INFRASTRUCTURE (CPU tier): stand-in for <cooperative_groups.h>. A cooperative kernel is emulated with ONE CTA (the
// grid barrier is then the CTA barrier) or, with -DMOF_EMUL_THREADS, with one OS thread per CTA (emul_runtime.cpp).
#pragma once
#include "emul_cuda_runtime.h"

namespace cooperative_groups {
struct grid_group {
    void sync() const { mof_emul::grid_sync(); }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
