// TEST INFRASTRUCTURE: the fiber scheduler behind emul_cuda_runtime.h. One OS thread; every thread of the block being
// "executed" is a ucontext fiber; fibers run round-robin and give the processor back at synchronisation points:
//   __syncthreads()   a counted barrier over the block's live threads (a thread that left the kernel stops counting)
//   __shfl_*_sync()   publish the value, barrier over the live lanes of the WARP, read the source lane, barrier again
// Blocks run one after the other (so `static` stands in for __shared__), atomics are plain read-modify-writes.
#include "emul_cuda_runtime.h"

EmulDim blockIdx, blockDim, threadIdx, gridDim;

namespace {
constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 128 * 1024;
ucontext_t mainCtx, fiberCtx[kMaxThreads];
char* stacks[kMaxThreads];
bool finished[kMaxThreads];
int current = -1, blockThreads = 0;
const std::function<void()>* body = nullptr;

int live = 0, arrived = 0;
unsigned long long generation = 0;
int warpLive[kMaxThreads / 32], warpArrived[kMaxThreads / 32];
unsigned long long warpGeneration[kMaxThreads / 32];
unsigned long long slots[kMaxThreads];  // shuffle exchange, 8 bytes per thread

void yield() { swapcontext(&fiberCtx[current], &mainCtx); }

void fiber_entry() {
    (*body)();
    const int me = current;
    finished[me] = true;
    // a thread that has left no longer takes part in barriers: release whoever is waiting for it
    live--, warpLive[me >> 5]--;
    if (live > 0 && arrived == live) arrived = 0, generation++;
    if (warpLive[me >> 5] > 0 && warpArrived[me >> 5] == warpLive[me >> 5]) warpArrived[me >> 5] = 0, warpGeneration[me >> 5]++;
    swapcontext(&fiberCtx[me], &mainCtx);
}

void warp_barrier() {
    const int w = current >> 5;
    const unsigned long long mine = warpGeneration[w];
    if (++warpArrived[w] == warpLive[w]) warpArrived[w] = 0, warpGeneration[w]++;
    while (warpGeneration[w] == mine) yield();
}
}  // namespace

void __syncthreads() {
    const unsigned long long mine = generation;
    if (++arrived == live) arrived = 0, generation++;
    while (generation == mine) yield();
}

namespace mof_emul {

unsigned long long shuffle(unsigned long long bits, int srcLane) {
    const int me = current, w = me >> 5;
    slots[me] = bits;
    warp_barrier();
    const int src = 32 * w + srcLane;
    const unsigned long long got = (srcLane >= 0 && srcLane < 32 && src < blockThreads && !finished[src]) ? slots[src] : bits;
    warp_barrier();
    return got;
}
int lane() { return current & 31; }

void launch(long long grid, int block, const std::function<void()>& b) {
    body = &b;
    blockThreads = block;
    blockDim.x = (unsigned)block, gridDim.x = (unsigned)grid;
    for (int t = 0; t < block; t++)
        if (!stacks[t]) stacks[t] = (char*)malloc(kStack);
    for (long long bi = 0; bi < grid; bi++) {
        blockIdx.x = (unsigned)bi;
        live = block, arrived = 0;
        for (int w = 0; w < (block + 31) / 32; w++) warpLive[w] = std::min(32, block - 32 * w), warpArrived[w] = 0;
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
