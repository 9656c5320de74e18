// TEST INFRASTRUCTURE: the fiber scheduler behind emul_cuda_runtime.h. One OS thread; the threads of the block being
// "executed" run on ucontext fibers; a thread that has to wait gives the processor back at synchronisation points:
//   __syncthreads()   a counted barrier over the block's live threads (a thread that left the kernel stops counting)
//   __shfl_*_sync()   publish the value, barrier over the live lanes of the WARP, read the source lane, barrier again
// Blocks run one after the other (so `static` stands in for __shared__), atomics are plain read-modify-writes.
#include "emul_cuda_runtime.h"

#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>

MOF_EMUL_TLS EmulDim blockIdx, blockDim, threadIdx, gridDim;

// Context switch. x86-64: a dozen instructions (callee-saved registers and the stack pointer; every fiber shares the
// floating-point control state) — glibc's swapcontext makes a system call per switch (the signal mask), and a kernel
// with barriers switches some twenty times per thread. Elsewhere, or with -DMOF_EMUL_UCONTEXT: ucontext.
#if defined(__x86_64__) && !defined(MOF_EMUL_UCONTEXT)
#define MOF_EMUL_ASM_SWITCH 1
asm(R"(
.text
.globl mof_emul_switch
.type mof_emul_switch,@function
mof_emul_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size mof_emul_switch,.-mof_emul_switch
)");
extern "C" void mof_emul_switch(void** saveSp, void* loadSp);
#endif

namespace {
constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 128 * 1024;
// A fiber runs thread after thread of the block (the next one that has not started yet) for as long as they run to
// completion; the first of its threads that has to wait at a synchronisation point parks on it. A kernel without
// barriers or shuffles therefore runs on ONE fiber with no context switch at all.
MOF_EMUL_TLS char* stacks[kMaxThreads];
MOF_EMUL_TLS bool finished[kMaxThreads];
MOF_EMUL_TLS int fiberOf[kMaxThreads];
MOF_EMUL_TLS int current = -1, blockThreads = 0, nextToStart = 0;
MOF_EMUL_TLS const std::function<void()>* body = nullptr;

MOF_EMUL_TLS int live = 0, arrived = 0;
MOF_EMUL_TLS unsigned long long generation = 0;
MOF_EMUL_TLS int warpLive[kMaxThreads / 32], warpArrived[kMaxThreads / 32];
MOF_EMUL_TLS unsigned long long warpGeneration[kMaxThreads / 32];
MOF_EMUL_TLS unsigned long long slots[kMaxThreads];  // shuffle exchange, 8 bytes per thread

MOF_EMUL_TLS int runningFiber = -1;
void fiber_entry();
#ifdef MOF_EMUL_ASM_SWITCH
MOF_EMUL_TLS void* mainSp = nullptr;
MOF_EMUL_TLS void* fiberSp[kMaxThreads];
void fiber_init(int f) {
    // the frame mof_emul_switch unwinds on the first switch: six registers, then `ret` into fiber_entry with the stack
    // pointer where a call would have left it (8 below a 16-byte boundary)
    uintptr_t top = ((uintptr_t)stacks[f] + kStack) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;               // where fiber_entry's caller's return address would be
    *--sp = (void*)&fiber_entry;
    for (int i = 0; i < 6; i++) *--sp = nullptr;
    fiberSp[f] = (void*)sp;
}
void to_fiber(int f) { mof_emul_switch(&mainSp, fiberSp[f]); }
void to_main(int f) { mof_emul_switch(&fiberSp[f], mainSp); }
#else
MOF_EMUL_TLS ucontext_t mainCtx, fiberCtx[kMaxThreads];
void fiber_init(int f) {
    getcontext(&fiberCtx[f]);
    fiberCtx[f].uc_stack.ss_sp = stacks[f], fiberCtx[f].uc_stack.ss_size = kStack, fiberCtx[f].uc_link = &mainCtx;
    makecontext(&fiberCtx[f], fiber_entry, 0);
}
void to_fiber(int f) { swapcontext(&mainCtx, &fiberCtx[f]); }
void to_main(int f) { swapcontext(&fiberCtx[f], &mainCtx); }
#endif

void yield() { to_main(fiberOf[current]); }

void fiber_entry() {
    const int f = runningFiber;
    while (nextToStart < blockThreads) {
        const int me = nextToStart++;
        current = me, threadIdx.x = (unsigned)me, fiberOf[me] = f;
        (*body)();
        finished[me] = true;
        // a thread that has left no longer takes part in barriers: release whoever is waiting for it
        live--, warpLive[me >> 5]--;
        if (live > 0 && arrived == live) arrived = 0, generation++;
        if (warpLive[me >> 5] > 0 && warpArrived[me >> 5] == warpLive[me >> 5]) warpArrived[me >> 5] = 0, warpGeneration[me >> 5]++;
    }
    for (;;) to_main(f);  // never resumed: the next block re-initialises the fiber
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

namespace {
MOF_EMUL_TLS EmulGraph* capture = nullptr;
}
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) {
    if (capture && !capture->isBody) delete capture;
    capture = new EmulGraph();
    return cudaSuccess;
}
cudaError_t cudaStreamBeginCaptureToGraph(cudaStream_t, cudaGraph_t graph, const cudaGraphNode_t*, const void*, size_t, cudaStreamCaptureMode) {
    capture = graph;  // owned by the graph of its conditional node
    return graph ? cudaSuccess : 1;
}
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* graph) {
    *graph = capture;
    capture = nullptr;
    return *graph ? cudaSuccess : 1;
}
cudaError_t cudaStreamGetCaptureInfo(cudaStream_t, cudaStreamCaptureStatus* status, unsigned long long* id, cudaGraph_t* graph, const cudaGraphNode_t** deps, size_t* ndeps) {
    if (status) *status = capture ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
    if (id) *id = 0;
    if (graph) *graph = capture;
    if (deps) *deps = nullptr;
    if (ndeps) *ndeps = 0;
    return cudaSuccess;
}

namespace mof_emul {

bool capturing() { return capture != nullptr; }
void record(std::function<void()> op) { capture->ops.push_back(std::move(op)); }
void submit(long long grid, int block, std::function<void()> b) {
    if (capture) capture->ops.push_back([grid, block, b] { launch(grid, block, b); });
    else launch(grid, block, b);
}
namespace {
MOF_EMUL_TLS std::vector<void*>* clusterSmem = nullptr;  // the dynamic shared memory of every CTA of the running cooperative launch
}
void* dynamic_smem(size_t bytes) {
    static MOF_EMUL_TLS char* buffer = nullptr;
    static MOF_EMUL_TLS size_t capacity = 0;
    if (bytes > capacity) {
        free(buffer);
        buffer = (char*)malloc(bytes);
        capacity = bytes;
    }
    if (clusterSmem && threadIdx.x == 0) (*clusterSmem)[blockIdx.x] = buffer;
    return buffer;
}
// Distributed shared memory: CTA `rank`'s buffer (valid after a grid_sync() that follows every CTA's dynamic_smem call).
void* peer_smem(void* mine, int rank) { return clusterSmem ? (*clusterSmem)[rank] : mine; }
void submit_cooperative(long long grid, int block, std::function<void()> b) {
    if (capture) capture->ops.push_back([grid, block, b] { launch_cooperative(grid, block, b); });
    else launch_cooperative(grid, block, b);
}

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

// One thread block of a grid, thread by thread, on the calling OS thread.
static void run_block(long long bi, long long grid, int block, const std::function<void()>& b) {
    body = &b;
    blockThreads = block;
    blockDim.x = (unsigned)block, gridDim.x = (unsigned)grid;
    blockIdx.x = (unsigned)bi;
    live = block, arrived = 0, nextToStart = 0;
    for (int w = 0; w < (block + 31) / 32; w++) warpLive[w] = std::min(32, block - 32 * w), warpArrived[w] = 0;
    for (int t = 0; t < block; t++) finished[t] = false, fiberOf[t] = -1;
    // start every thread: each fiber takes threads until one of them has to wait
    for (int f = 0; nextToStart < block; f++) {
        if (!stacks[f]) stacks[f] = (char*)malloc(kStack);
        fiber_init(f);
        runningFiber = f;
        to_fiber(f);
    }
    // then round-robin over the waiting ones
    for (bool any = true; any;) {
        any = false;
        for (int t = 0; t < block; t++) {
            if (finished[t]) continue;
            current = t, threadIdx.x = (unsigned)t;
            to_fiber(fiberOf[t]);
            any = any || !finished[t];
        }
    }
}

void launch(long long grid, int block, const std::function<void()>& b) {
    for (long long bi = 0; bi < grid; bi++) run_block(bi, grid, block, b);
}

// A cooperative launch: the CTAs have to be alive together (they meet in grid_sync), and a `static` stands in for
// __shared__ — so they can only be alive together on different OS threads with per-thread statics (-DMOF_EMUL_THREADS).
// Without it a cooperative kernel is run as ONE CTA (cudaDeviceGetAttribute reports one multiprocessor).
namespace {
struct GridBarrier {
    std::mutex m;
    std::condition_variable cv;
    int size = 0, waiting = 0;
    unsigned long long generation = 0;
    void wait() {
        std::unique_lock<std::mutex> lock(m);
        const unsigned long long mine = generation;
        if (++waiting == size) waiting = 0, generation++, cv.notify_all();
        else cv.wait(lock, [&] { return generation != mine; });
    }
};
MOF_EMUL_TLS GridBarrier* gridBarrier = nullptr;
}  // namespace

void grid_sync() {
    __syncthreads();
    if (gridBarrier && threadIdx.x == 0) gridBarrier->wait();
    __syncthreads();
}

void launch_cooperative(long long grid, int block, const std::function<void()>& b) {
#ifdef MOF_EMUL_THREADS
    if (grid > 1) {
        GridBarrier barrier;
        barrier.size = (int)grid;
        std::vector<void*> smemOf((size_t)grid, nullptr);
        std::vector<std::thread> ctas;
        auto cta = [&](long long bi, bool worker) {
            gridBarrier = &barrier;
            clusterSmem = &smemOf;
            run_block(bi, grid, block, b);
            gridBarrier = nullptr;
            clusterSmem = nullptr;
            if (worker)  // the worker's fiber stacks die with it
                for (int f = 0; f < kMaxThreads; f++) free(stacks[f]), stacks[f] = nullptr;
        };
        for (long long bi = 1; bi < grid; bi++) ctas.emplace_back(cta, bi, true);
        cta(0, false);
        for (std::thread& t : ctas) t.join();
        return;
    }
#endif
    if (grid != 1) abort();  // see above
    run_block(0, 1, block, b);
}
}  // namespace mof_emul
