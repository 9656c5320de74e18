// TEST INFRASTRUCTURE: the fiber scheduler behind emul_cuda_runtime.h.
#include "emul_cuda_runtime.h"

EmulDim blockIdx, blockDim, threadIdx, gridDim;

namespace {
constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 128 * 1024;
ucontext_t mainCtx, fiberCtx[kMaxThreads];
char* stacks[kMaxThreads];
bool finished[kMaxThreads];
int current = -1;
const std::function<void()>* body = nullptr;

void fiber_entry() {
    (*body)();
    finished[current] = true;
    swapcontext(&fiberCtx[current], &mainCtx);
}
}  // namespace

void __syncthreads() { swapcontext(&fiberCtx[current], &mainCtx); }

namespace mof_emul {
void launch(long long grid, int block, const std::function<void()>& b) {
    body = &b;
    blockDim.x = (unsigned)block, gridDim.x = (unsigned)grid;
    for (int t = 0; t < block; t++)
        if (!stacks[t]) stacks[t] = (char*)malloc(kStack);
    for (long long bi = 0; bi < grid; bi++) {
        blockIdx.x = (unsigned)bi;
        for (int t = 0; t < block; t++) {
            getcontext(&fiberCtx[t]);
            fiberCtx[t].uc_stack.ss_sp = stacks[t], fiberCtx[t].uc_stack.ss_size = kStack, fiberCtx[t].uc_link = &mainCtx;
            makecontext(&fiberCtx[t], fiber_entry, 0);
            finished[t] = false;
        }
        for (bool any = true; any;) {
            any = false;
            for (int t = 0; t < block; t++) {
                if (finished[t]) continue;
                current = t, threadIdx.x = (unsigned)t;
                swapcontext(&mainCtx, &fiberCtx[t]);
                any = any || !finished[t];
            }
        }
    }
}
}  // namespace mof_emul
