// TEST INFRASTRUCTURE (CPU tier): host stand-ins for the few CUDA runtime calls and for the kernel-launch syntax that
// meshopticalflow_b200/csrc uses, so that a .cu file can be compiled by g++ with -DMOF_HOST_EMULATION and its REAL
// source — kernels and host driver alike — run on the CPU. "Device" memory is malloc'd host memory, streams and events
// are no-ops (everything is synchronous; a stream capture records launches and copies, a graph launch replays them), and
// a launch runs the kernel body once per (block, thread) on fibers of one OS thread (emul_runtime.cpp), a thread that
// reaches __syncthreads() or a warp shuffle giving way to the others: barrier semantics hold, __shared__ variables (made
// `static`) are per block because blocks run one after the other. A cooperative kernel is run as ONE CTA.
//
// Build with -fno-gnu-unique: several emulated libraries live in one pytest process, and a `static` inside an inline or
// template kernel (every __shared__ variable of one) is otherwise a STB_GNU_UNIQUE symbol that the dynamic linker shares
// between ALL of them — including between the per-thread (MOF_EMUL_THREADS) and the plain variant of the same variable.
#pragma once

#include <ucontext.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include <algorithm>
using std::max;  // CUDA's global-namespace overloads
using std::min;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
// -DMOF_EMUL_THREADS: several "GPUs" in one process, one OS thread each (the partitioned-mesh path, nccl.h here): the
// emulator's state and the kernels' __shared__ variables are then per thread.
#ifdef MOF_EMUL_THREADS
#define MOF_EMUL_TLS thread_local
#else
#define MOF_EMUL_TLS
#endif
#define __shared__ static MOF_EMUL_TLS
#define __launch_bounds__(...)

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }
// "inter-process" memory handles between the emulated ranks (OS threads of one process): the handle is the pointer
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return *p ? cudaSuccess : 1; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
namespace mof_emul {
// Stream capture: while a capture is open, launches and stream-ordered copies are RECORDED (with their arguments, by
// value) instead of run, like on the device; cudaGraphLaunch runs the recorded list.
bool capturing();
void record(std::function<void()> op);
}  // namespace mof_emul
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) {
    if (mof_emul::capturing()) mof_emul::record([=] { memset(p, v, bytes); });
    else memset(p, v, bytes);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t bytes, cudaMemcpyKind, cudaStream_t) {
    if (mof_emul::capturing()) mof_emul::record([=] { memmove(d, s, bytes); });
    else memmove(d, s, bytes);
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = *greatest = 0; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return cudaSuccess; }
typedef void* cudaMemPool_t;
enum { cudaMemPoolAttrReleaseThreshold = 0 };
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) { *pool = nullptr; return cudaSuccess; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void*) { return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t bytes, cudaMemcpyKind) { memmove(d, s, bytes); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
// cooperative launches: ONE CTA (a `static` stands in for __shared__, so CTAs cannot be alive together on one OS thread), or with
// -DMOF_EMUL_THREADS three CTAs on three OS threads (emul_runtime.cpp: launch_cooperative)
template <class K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 1; return cudaSuccess; }
enum { cudaDevAttrMultiProcessorCount = 16 };
#ifdef MOF_EMUL_THREADS
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 3; return cudaSuccess; }  // three "multiprocessors": three CTAs, three OS threads
#else
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return cudaSuccess; }
#endif
enum { cudaHostAllocDefault = 0 };
inline cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }

// CUDA graphs by stream capture
struct EmulGraph {
    std::vector<std::function<void()>> ops;
    // conditional nodes: the graph owns the condition words of its handles (with their defaults, assigned at every launch)
    // and the body graphs of its WHILE nodes
    std::vector<std::pair<unsigned*, unsigned>> conditions;
    std::vector<EmulGraph*> bodies;
    bool isBody = false;
};
typedef EmulGraph* cudaGraph_t;
typedef EmulGraph* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal, cudaStreamCaptureModeRelaxed };
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode);
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* graph);
// (an instantiated graph shares the condition words and bodies of the graph it came from: keep that one alive)
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* exec, cudaGraph_t graph, unsigned long long) { *exec = new EmulGraph(*graph); return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t) {
    for (auto& c : exec->conditions) *c.first = c.second;
    for (auto& op : exec->ops) op();
    return cudaSuccess;
}
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) {
    if (g) {
        for (auto& c : g->conditions) delete c.first;
        for (EmulGraph* b : g->bodies) cudaGraphDestroy(b);
    }
    delete g;
    return cudaSuccess;
}
// Conditional WHILE nodes, as far as multigrid.cu uses them: a handle is a pointer to the graph's condition word, the node an
// operation that runs its body graph while the word is non-zero, the body filled by a capture "to graph".
typedef unsigned* cudaGraphConditionalHandle;
typedef void* cudaGraphNode_t;
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone, cudaStreamCaptureStatusActive };
enum { cudaGraphCondAssignDefault = 1, cudaStreamSetCaptureDependencies = 1 };
enum cudaGraphNodeType { cudaGraphNodeTypeConditional = 13 };
enum cudaGraphConditionalNodeType { cudaGraphCondTypeIf = 0, cudaGraphCondTypeWhile = 1 };
struct cudaConditionalNodeParams {
    cudaGraphConditionalHandle handle;
    cudaGraphConditionalNodeType type;
    unsigned size;
    cudaGraph_t* phGraph_out;
};
struct cudaGraphNodeParams {
    cudaGraphNodeType type;
    cudaConditionalNodeParams conditional;
};
cudaError_t cudaStreamGetCaptureInfo(cudaStream_t, cudaStreamCaptureStatus* status, unsigned long long* id, cudaGraph_t* graph, const cudaGraphNode_t** deps, size_t* ndeps);
inline cudaError_t cudaGraphConditionalHandleCreate(cudaGraphConditionalHandle* h, cudaGraph_t graph, unsigned defaultValue, unsigned) {
    *h = new unsigned(defaultValue);
    graph->conditions.push_back({*h, defaultValue});
    return cudaSuccess;
}
inline cudaError_t cudaGraphAddNode(cudaGraphNode_t* node, cudaGraph_t graph, const cudaGraphNode_t*, size_t, cudaGraphNodeParams* params) {
    if (params->type != cudaGraphNodeTypeConditional || params->conditional.type != cudaGraphCondTypeWhile) return 1;
    EmulGraph* body = new EmulGraph();
    body->isBody = true;
    graph->bodies.push_back(body);
    unsigned* cond = params->conditional.handle;
    graph->ops.push_back([body, cond] {
        while (*cond)
            for (auto& op : body->ops) op();
    });
    params->conditional.phGraph_out = &graph->bodies.back();
    *node = body;
    return cudaSuccess;
}
inline cudaError_t cudaStreamUpdateCaptureDependencies(cudaStream_t, cudaGraphNode_t*, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaStreamBeginCaptureToGraph(cudaStream_t, cudaGraph_t graph, const cudaGraphNode_t*, const void*, size_t, cudaStreamCaptureMode);
inline void cudaGraphSetConditional(cudaGraphConditionalHandle h, unsigned value) { *h = value; }

// vector types and cache-hinted loads
struct alignas(8) float2 { float x, y; };     // the device's alignment requirements: -fsanitize=alignment then reports a vector
struct alignas(16) double2 { double x, y; };  // load that would be a "misaligned address" fault on the GPU
struct alignas(16) float4 { float x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }

struct EmulDim { unsigned x = 1, y = 1, z = 1; };
extern MOF_EMUL_TLS EmulDim blockIdx, blockDim, threadIdx, gridDim;
void __syncthreads();

namespace mof_emul {
void launch(long long grid, int block, const std::function<void()>& body);
// A kernel launch as the stream sees it: run now, or recorded into the capture in progress. `body` owns its arguments.
void submit(long long grid, int block, std::function<void()> body);
// Cooperative kernels: all CTAs alive together, one OS thread each (MOF_EMUL_THREADS), meeting in grid_sync(); else one CTA.
void launch_cooperative(long long grid, int block, const std::function<void()>& body);
void submit_cooperative(long long grid, int block, std::function<void()> body);  // ... as the stream sees it (run now, or recorded)
void grid_sync();
void* dynamic_smem(size_t bytes);  // the running CTA's dynamic shared memory (one buffer per OS thread, grown on demand)
void* peer_smem(void* mine, int rank);  // ... and that of another CTA of the same cooperative launch (threads build; else `mine`)
unsigned long long shuffle(unsigned long long bits, int srcLane);  // warp-synchronous exchange of 8 bytes
int lane();
template <class T>
inline T shuffle_value(T v, int srcLane) {
    static_assert(sizeof(T) <= 8, "shuffle of up to 8 bytes");
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    bits = shuffle(bits, srcLane);
    memcpy(&v, &bits, sizeof(T));
    return v;
}
}  // namespace mof_emul

// warp shuffles (full-warp participation of the live lanes is assumed, as in the kernels of this repository)
template <class T> inline T __shfl_sync(unsigned, T v, int srcLane, int = 32) { return mof_emul::shuffle_value(v, srcLane); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return mof_emul::shuffle_value(v, mof_emul::lane() - (int)d); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return mof_emul::shuffle_value(v, mof_emul::lane() + (int)d); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return mof_emul::shuffle_value(v, mof_emul::lane() ^ m); }
// atomics: one OS thread, fibers switch only at synchronisation points
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicCAS(T* p, T expected, T desired) { T o = *p; if (o == expected) *p = desired; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
inline void __threadfence() {}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __syncwarp(unsigned = 0xffffffffu) {}
