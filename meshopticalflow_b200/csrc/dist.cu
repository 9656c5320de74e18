// One mesh over several GPUs (BASELINE.json configs[4]): the rows of the two families of linear systems
// are split into contiguous blocks, one block per rank (one process per GPU) — the E edge unknowns of the flow system
// in blocks of 32-row slices, the V vertices of the smoothing systems in equal ranges. Everything else — mesh
// operators, signals, walks, the coarse multigrid levels — is replicated: every rank makes the same calls with the same
// inputs and holds full-length vectors, of which it computes its own rows. What crosses NVLink, through NCCL on the
// context's stream:
//   * halo exchange before every fine-level SpMV: the entries of the input vector that a rank's rows reference outside
//     its block (index lists built once per mesh from the matrix patterns, grouped ncclSend/ncclRecv of packed values),
//   * all-reduce of the PCG dot products and of the level-1 restriction (each rank restricts its own rows),
//   * one all-gather of the solution at the end of a solve.
// The communicator is created from an id that the host side broadcasts (torch.distributed does that in bench.py and
// the tests); this library never touches the rendezvous. world == 1 runs the same code with empty exchanges.
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "mof_internal.cuh"

namespace mof {

// Row blocks and halo index lists of one system. kind 0 = FLOW (one value per edge), 1 = SCALAR (six per vertex).
struct Partition {
    int width = 1;
    std::vector<int> rowStart;                                  // world + 1
    std::vector<int> sendCount, sendOff, recvCount, recvOff;    // per peer, in indices
    int nSend = 0, nRecv = 0;
    DBuf<int> sendIdx, recvIdx;
    DBuf<int> offs;  // device copy for the peer-memory exchange: sendOff[0 .. world], then recvOff[0 .. world]
};

// Peer-memory halo exchange (MOF_DIST_P2P=1; off by default, see dist_p2p_setup): every rank owns a WINDOW of device
// memory (cudaMalloc, shared through cudaIpc handles) that its peers write into directly over NVLink — one region per (source rank,
// slot) plus one arrival flag each. An exchange is two kernels and no library call: k_halo_put packs the values a peer needs straight
// into that peer's window and, from its last CTA, raises the flag with the pair's sequence number; k_halo_wait polls the flags of the
// ranks it expects data from and unpacks. Two slots per pair, used alternately, are enough: the next message to a peer is only sent
// after that peer's message of the current exchange has arrived, which it sends before it unpacks — so it can be at most one message
// behind (pairs exchange in both directions or not at all: an empty message still raises the flag). Sequence numbers live in device
// memory, so the two kernels can sit in a replayed CUDA graph.
struct PeerWindow {
    bool on = false;
    bool shared = false;                   // every rank has mapped every window (agreed): releasing them is a collective step
    unsigned char* base = nullptr;
    size_t bytes = 0, flagBytes = 0, haloCap = 0;
    std::vector<unsigned char*> peer;      // peer[j]: rank j's window mapped into this process (peer[rank] = base)
    DBuf<unsigned char*> table;            // the same on the device
    DBuf<unsigned long long> seq;          // [0, w) messages sent to j; [w, 2w) received from j; [2w], [2w+1] CTA counters; [2w+2] time-out flag
};

struct DistState {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    bool meshReady = false;
    std::vector<int> sliceStart;            // FLOW: world + 1 (32-row slices)
    Partition part[2];
    std::vector<Partition*> extra;          // further partitions made by the solvers (multigrid levels dealt in cell ranges): dist_add_partition
    DBuf<double> sendBuf, recvBuf;          // sized for the widest exchange in fp64, reused for fp32
    PeerWindow win;
    // MOF_DIST_TRACE=1 (with MOF_DIST_GRAPH=0): device time and count per kind of exchange, printed by rank 0 when the context closes
    bool trace = false;
    double traceMs[4] = {0, 0, 0, 0};
    long long traceN[4] = {0, 0, 0, 0};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
};
enum { TR_HALO = 0, TR_ALLREDUCE = 1, TR_ALLGATHER = 2, TR_HALO_BYTES = 3 };
struct TraceScope {
    mof_ctx* ctx;
    DistState& d;
    int cat;
    TraceScope(mof_ctx* c, int category) : ctx(c), d(*c->dist), cat(category) {
        if (d.trace) cudaEventRecord(d.t0, ctx->stream);
    }
    ~TraceScope() {
        if (!d.trace) return;
        cudaEventRecord(d.t1, ctx->stream);
        cudaEventSynchronize(d.t1);
        float ms = 0;
        cudaEventElapsedTime(&ms, d.t0, d.t1);
        d.traceMs[cat] += ms, d.traceN[cat]++;
    }
};

namespace {

#define MOF_NCCL(call)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) return fail(ctx, MOF_E_CUDA, std::string(#call) + ": " + ncclGetErrorString(r__)); \
    } while (0)

constexpr int B = 256;

// flags[c] = 1 for every column outside [r0, r1) referenced by the slices [s0, s1) of the sliced FLOW pattern ...
__global__ void k_mark_halo_sell(const int* __restrict__ sliceBase, const int* __restrict__ col, int s0, int s1, int r0, int r1, int* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int s = s0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < s1; s += warps) {
        const int base = sliceBase[s];
        const int len = (sliceBase[s + 1] - base) >> 5;
        for (int j = 0; j < len; j++) {
            int c = col[(size_t)base + 32 * (size_t)j + lane];
            if (c < r0 || c >= r1) flags[c] = 1;
        }
    }
}
// ... and by the rows [r0, r1) of the SCALAR CSR pattern.
__global__ void k_mark_halo_csr(const int* __restrict__ rowptr, const int* __restrict__ col, int r0, int r1, int* __restrict__ flags) {
    int v = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= r1) return;
    for (int k = rowptr[v]; k < rowptr[v + 1]; k++) {
        int c = col[k];
        if (c < r0 || c >= r1) flags[c] = 1;
    }
}
__global__ void k_compact(const int* __restrict__ flags, const int* __restrict__ pos, int n, int* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[pos[i]] = i;
}
template <class T>
__global__ void k_pack(const T* __restrict__ vec, const int* __restrict__ idx, int n, int width, T* __restrict__ buf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * width) return;
    int e = i / width, c = i - e * width;
    buf[i] = vec[(size_t)idx[e] * width + c];
}
template <class T>
__global__ void k_unpack(const T* __restrict__ buf, const int* __restrict__ idx, int n, int width, T* __restrict__ vec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * width) return;
    int e = i / width, c = i - e * width;
    vec[(size_t)idx[e] * width + c] = buf[i];
}

#ifdef MOF_HOST_EMULATION
inline unsigned long long p2p_clock() { return 0; }
template <class T> inline T p2p_load(const T* p) { return *(const volatile T*)p; }
#else
__device__ __forceinline__ unsigned long long p2p_clock() { return (unsigned long long)clock64(); }
template <class T> __device__ __forceinline__ T p2p_load(const T* p) { return __ldcv(p); }  // written by another GPU: never from L1
#endif
constexpr unsigned long long P2P_SPIN_LIMIT = 20000000000ull;  // ~10 s of SM clock: a peer that never arrives is an error, not a hang

// offs: sendOff[0 .. world], recvOff[0 .. world]. A pair is active in an exchange when either direction carries entries.
__device__ __forceinline__ bool p2p_active(const int* offs, int world, int j) { return offs[j + 1] > offs[j] || offs[world + 1 + j + 1] > offs[world + 1 + j]; }

template <class T>
__global__ void k_halo_put(const T* __restrict__ vec, const int* __restrict__ sendIdx, const int* __restrict__ offs, int world, int me, int width,
                           unsigned char* const* __restrict__ peer, size_t flagBytes, size_t haloCap, unsigned long long* seq) {
    const long long total = (long long)offs[world] * width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / width), c = (int)(i - (long long)e * width);
        int j = 0;
        while (e >= offs[j + 1]) j++;
        const unsigned long long s = seq[j] + 1;  // (raised by the last CTA below, after every CTA has read it)
        T* dst = reinterpret_cast<T*>(peer[j] + flagBytes + ((size_t)me * 2 + (size_t)(s & 1)) * haloCap) + (size_t)(e - offs[j]) * width + c;
        *dst = vec[(size_t)sendIdx[e] * width + c];
    }
    __threadfence_system();
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&seq[2 * world], 1ull) == (unsigned long long)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    for (int j = threadIdx.x; j < world; j += blockDim.x)
        if (j != me && p2p_active(offs, world, j)) {
            const unsigned long long s = seq[j] + 1;
            seq[j] = s;
            volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(peer[j]) + ((size_t)me * 2 + (size_t)(s & 1));
            *flag = s;
        }
    if (threadIdx.x == 0) seq[2 * world] = 0;
}

template <class T>
__global__ void k_halo_wait(T* __restrict__ vec, const int* __restrict__ recvIdx, const int* __restrict__ offs, int world, int me, int width,
                            const unsigned char* __restrict__ mine, size_t flagBytes, size_t haloCap, unsigned long long* seq) {
    const int* recvOff = offs + world + 1;
    if ((int)threadIdx.x < world) {
        const int j = threadIdx.x;
        if (j != me && p2p_active(offs, world, j)) {
            const unsigned long long s = seq[world + j] + 1;
            const volatile unsigned long long* flag = reinterpret_cast<const volatile unsigned long long*>(mine) + ((size_t)j * 2 + (size_t)(s & 1));
            const unsigned long long t0 = p2p_clock();
            while (*flag < s)
                if (p2p_clock() - t0 > P2P_SPIN_LIMIT) {
                    seq[2 * world + 2] = 1;
                    break;
                }
        }
    }
    __syncthreads();
    __threadfence_system();
    const long long total = (long long)recvOff[world] * width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / width), c = (int)(i - (long long)e * width);
        int j = 0;
        while (e >= recvOff[j + 1]) j++;
        const unsigned long long s = seq[world + j] + 1;
        const T* src = reinterpret_cast<const T*>(mine + flagBytes + ((size_t)j * 2 + (size_t)(s & 1)) * haloCap) + (size_t)(e - recvOff[j]) * width + c;
        vec[(size_t)recvIdx[e] * width + c] = p2p_load(src);
    }
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&seq[2 * world + 1], 1ull) == (unsigned long long)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    for (int j = threadIdx.x; j < world; j += blockDim.x)
        if (j != me && p2p_active(offs, world, j)) seq[world + j] += 1;
    if (threadIdx.x == 0) seq[2 * world + 1] = 0;
}

// The two as ONE launch: a CTA writes its share of the outgoing entries, the last one to finish raises the flags, and every CTA goes on to
// poll for the peers' messages and unpack its share of them. Waiting depends on the PEER's writes, not on this grid's, and the grid is small
// enough to be resident as a whole (<= 2 CTAs per SM), so no CTA ever waits for one that cannot run.
template <class T>
__global__ void k_halo_exchange(T* __restrict__ vec, const int* __restrict__ sendIdx, const int* __restrict__ recvIdx, const int* __restrict__ offs, int world, int me,
                                int width, unsigned char* const* __restrict__ peer, const unsigned char* __restrict__ mine, size_t flagBytes, size_t haloCap,
                                unsigned long long* seq) {
    __shared__ int last;
    {
        const long long total = (long long)offs[world] * width;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int e = (int)(i / width), c = (int)(i - (long long)e * width);
            int j = 0;
            while (e >= offs[j + 1]) j++;
            const unsigned long long s = seq[j] + 1;
            T* dst = reinterpret_cast<T*>(peer[j] + flagBytes + ((size_t)me * 2 + (size_t)(s & 1)) * haloCap) + (size_t)(e - offs[j]) * width + c;
            *dst = vec[(size_t)sendIdx[e] * width + c];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) last = atomicAdd(&seq[2 * world], 1ull) == (unsigned long long)gridDim.x - 1;
        __syncthreads();
        if (last) {
            __threadfence_system();
            for (int j = threadIdx.x; j < world; j += blockDim.x)
                if (j != me && p2p_active(offs, world, j)) {
                    const unsigned long long s = seq[j] + 1;
                    seq[j] = s;
                    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(peer[j]) + ((size_t)me * 2 + (size_t)(s & 1));
                    *flag = s;
                }
            if (threadIdx.x == 0) seq[2 * world] = 0;
        }
        __syncthreads();
    }
    const int* recvOff = offs + world + 1;
    if ((int)threadIdx.x < world) {
        const int j = threadIdx.x;
        if (j != me && p2p_active(offs, world, j)) {
            const unsigned long long s = seq[world + j] + 1;
            const volatile unsigned long long* flag = reinterpret_cast<const volatile unsigned long long*>(mine) + ((size_t)j * 2 + (size_t)(s & 1));
            const unsigned long long t0 = p2p_clock();
            while (*flag < s)
                if (p2p_clock() - t0 > P2P_SPIN_LIMIT) {
                    seq[2 * world + 2] = 1;
                    break;
                }
        }
    }
    __syncthreads();
    __threadfence_system();
    const long long total = (long long)recvOff[world] * width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / width), c = (int)(i - (long long)e * width);
        int j = 0;
        while (e >= recvOff[j + 1]) j++;
        const unsigned long long s = seq[world + j] + 1;
        const T* src = reinterpret_cast<const T*>(mine + flagBytes + ((size_t)j * 2 + (size_t)(s & 1)) * haloCap) + (size_t)(e - recvOff[j]) * width + c;
        vec[(size_t)recvIdx[e] * width + c] = p2p_load(src);
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&seq[2 * world + 1], 1ull) == (unsigned long long)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    for (int j = threadIdx.x; j < world; j += blockDim.x)
        if (j != me && p2p_active(offs, world, j)) seq[world + j] += 1;
    if (threadIdx.x == 0) seq[2 * world + 1] = 0;
}

// All-gather of every rank's own range of up to two fp32 vectors through the same windows: a message to EVERY peer (so every pair is
// active), laid out [vector][own range]; the receiver copies each peer's range into place.
struct GatherVecs {
    float* v[2];
    int count;
};
__global__ void k_gather_put(GatherVecs vs, const int* __restrict__ rowStart, int world, int me, int width, unsigned char* const* __restrict__ peer, size_t flagBytes,
                             size_t haloCap, unsigned long long* seq) {
    const long long own = (long long)(rowStart[me + 1] - rowStart[me]) * width, first = (long long)rowStart[me] * width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < own * vs.count; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i / own);
        const float x = vs.v[v][first + (i - (long long)v * own)];
        for (int j = 0; j < world; j++) {
            if (j == me) continue;
            const unsigned long long s = seq[j] + 1;
            reinterpret_cast<float*>(peer[j] + flagBytes + ((size_t)me * 2 + (size_t)(s & 1)) * haloCap)[i] = x;
        }
    }
    __threadfence_system();
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&seq[2 * world], 1ull) == (unsigned long long)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    for (int j = threadIdx.x; j < world; j += blockDim.x)
        if (j != me) {
            const unsigned long long s = seq[j] + 1;
            seq[j] = s;
            volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(peer[j]) + ((size_t)me * 2 + (size_t)(s & 1));
            *flag = s;
        }
    if (threadIdx.x == 0) seq[2 * world] = 0;
}
__global__ void k_gather_wait(GatherVecs vs, const int* __restrict__ rowStart, int world, int me, int width, const unsigned char* __restrict__ mine, size_t flagBytes,
                              size_t haloCap, unsigned long long* seq) {
    if ((int)threadIdx.x < world && (int)threadIdx.x != me) {
        const int j = threadIdx.x;
        const unsigned long long s = seq[world + j] + 1;
        const volatile unsigned long long* flag = reinterpret_cast<const volatile unsigned long long*>(mine) + ((size_t)j * 2 + (size_t)(s & 1));
        const unsigned long long t0 = p2p_clock();
        while (*flag < s)
            if (p2p_clock() - t0 > P2P_SPIN_LIMIT) {
                seq[2 * world + 2] = 1;
                break;
            }
    }
    __syncthreads();
    __threadfence_system();
    // element i of the concatenation of the peers' messages (my own range is already in place)
    const long long all = (long long)rowStart[world] * width, mineLen = (long long)(rowStart[me + 1] - rowStart[me]) * width;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (all - mineLen) * vs.count; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i / (all - mineLen));
        long long r = i - (long long)v * (all - mineLen);  // position among the rows that are not mine
        if (r >= (long long)rowStart[me] * width) r += mineLen;
        int j = 0;
        while (r >= (long long)rowStart[j + 1] * width) j++;
        const long long len = (long long)(rowStart[j + 1] - rowStart[j]) * width, at = r - (long long)rowStart[j] * width;
        const unsigned long long s = seq[world + j] + 1;
        const float* src = reinterpret_cast<const float*>(mine + flagBytes + ((size_t)j * 2 + (size_t)(s & 1)) * haloCap) + (long long)v * len + at;
        vs.v[v][r] = p2p_load(src);
    }
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&seq[2 * world + 1], 1ull) == (unsigned long long)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    for (int j = threadIdx.x; j < world; j += blockDim.x)
        if (j != me) seq[world + j] += 1;
    if (threadIdx.x == 0) seq[2 * world + 1] = 0;
}

template <class T>
int halo_exchange(mof_ctx* ctx, Partition& p, T* vec, ncclDataType_t type) {
    DistState& d = *ctx->dist;
    if (d.world == 1) return MOF_OK;
    TraceScope ts(ctx, TR_HALO);
    if (d.trace) d.traceMs[TR_HALO_BYTES] += (double)(p.nSend + p.nRecv) * p.width * sizeof(T), d.traceN[TR_HALO_BYTES]++;
    const int w = p.width;
    if (d.win.on) {
        const PeerWindow& win = d.win;
        static const bool fused = !(getenv("MOF_DIST_P2P_FUSED") && *getenv("MOF_DIST_P2P_FUSED") == '0');
        const int putGrid = std::max(1, std::min(kSMs * 2, blocks_for((long long)p.nSend * w, B)));
        const int waitGrid = std::max(1, std::min(kSMs * 2, blocks_for((long long)p.nRecv * w, B)));
        if (fused) {
#ifdef MOF_HOST_EMULATION
            const int grid = 1;  // (the emulator runs CTAs one after the other: a CTA polling for a peer would keep the flag-raising one from running)
#else
            const int grid = std::max(putGrid, waitGrid);
#endif
            MOF_LAUNCH(k_halo_exchange<T>, grid, B, 0, vec, (const int*)p.sendIdx.p, (const int*)p.recvIdx.p, (const int*)p.offs.p, d.world, d.rank, w,
                       (unsigned char* const*)win.table.p, (const unsigned char*)win.base, win.flagBytes, win.haloCap, win.seq.p);
            return MOF_OK;
        }
        MOF_LAUNCH(k_halo_put<T>, putGrid, B, 0, (const T*)vec, (const int*)p.sendIdx.p, (const int*)p.offs.p, d.world, d.rank, w, (unsigned char* const*)win.table.p,
                   win.flagBytes, win.haloCap, win.seq.p);
        MOF_LAUNCH(k_halo_wait<T>, waitGrid, B, 0, vec, (const int*)p.recvIdx.p, (const int*)p.offs.p, d.world, d.rank, w, (const unsigned char*)win.base,
                   win.flagBytes, win.haloCap, win.seq.p);
        return MOF_OK;
    }
    T* sb = (T*)d.sendBuf.p;
    T* rb = (T*)d.recvBuf.p;
    if (p.nSend) MOF_LAUNCH(k_pack<T>, blocks_for((long long)p.nSend * w, B), B, 0, vec, p.sendIdx.p, p.nSend, w, sb);
    MOF_NCCL(ncclGroupStart());
    for (int j = 0; j < d.world; j++) {
        if (j == d.rank) continue;
        if (p.sendCount[j]) MOF_NCCL(ncclSend(sb + (size_t)p.sendOff[j] * w, (size_t)p.sendCount[j] * w, type, j, d.comm, ctx->stream));
        if (p.recvCount[j]) MOF_NCCL(ncclRecv(rb + (size_t)p.recvOff[j] * w, (size_t)p.recvCount[j] * w, type, j, d.comm, ctx->stream));
    }
    MOF_NCCL(ncclGroupEnd());
    if (p.nRecv) MOF_LAUNCH(k_unpack<T>, blocks_for((long long)p.nRecv * w, B), B, 0, rb, p.recvIdx.p, p.nRecv, w, vec);
    return MOF_OK;
}

// Halo lists of one partition whose out-of-block columns have been flagged in ctx->itmp0[0..n).
int build_lists(mof_ctx* ctx, Partition& p, int n) {
    DistState& d = *ctx->dist;
    const int N = d.world;
    DBuf<int>& flags = ctx->itmp0;
    DBuf<int>& pos = ctx->itmp1;
    MOF_TRY(exclusive_scan_int(ctx, flags.p, pos.p, n + 1, nullptr));
    int H = 0;
    MOF_CUDA(read_back(ctx, &H, pos.p + n));
    p.nRecv = H;
    MOF_CUDA(p.recvIdx.alloc((size_t)std::max(H, 1)));
    if (H) MOF_LAUNCH(k_compact, blocks_for(n, B), B, 0, flags.p, pos.p, n, p.recvIdx.p);
    std::vector<int> hIdx((size_t)H);
    if (H) MOF_CUDA(read_back(ctx, hIdx.data(), p.recvIdx.p, (size_t)H));
    p.sendCount.assign(N, 0), p.sendOff.assign(N, 0), p.recvCount.assign(N, 0), p.recvOff.assign(N, 0);
    for (int k = 0, i = 0; k < N; k++) {  // ascending indices: grouped by owner
        p.recvOff[k] = i;
        while (i < H && hIdx[i] < p.rowStart[k + 1]) i++;
        p.recvCount[k] = i - p.recvOff[k];
    }
    // who needs how much from whom: row k of the matrix = recvCount of rank k
    DBuf<int> counts, matrix;
    MOF_CUDA(counts.alloc(N));
    MOF_CUDA(matrix.alloc((size_t)N * N));
    MOF_CUDA(cudaMemcpyAsync(counts.p, p.recvCount.data(), sizeof(int) * N, cudaMemcpyHostToDevice, ctx->stream));
    MOF_NCCL(ncclAllGather(counts.p, matrix.p, N, ncclInt, d.comm, ctx->stream));
    std::vector<int> hm((size_t)N * N);
    MOF_CUDA(read_back(ctx, hm.data(), matrix.p, (size_t)N * N));
    counts.release(), matrix.release();
    p.nSend = 0;
    for (int j = 0; j < N; j++) {
        p.sendCount[j] = j == d.rank ? 0 : hm[(size_t)j * N + d.rank];
        p.sendOff[j] = p.nSend;
        p.nSend += p.sendCount[j];
    }
    MOF_CUDA(p.sendIdx.alloc((size_t)std::max(p.nSend, 1)));
    MOF_NCCL(ncclGroupStart());
    for (int j = 0; j < N; j++) {
        if (j == d.rank) continue;
        if (p.recvCount[j]) MOF_NCCL(ncclSend(p.recvIdx.p + p.recvOff[j], (size_t)p.recvCount[j], ncclInt, j, d.comm, ctx->stream));
        if (p.sendCount[j]) MOF_NCCL(ncclRecv(p.sendIdx.p + p.sendOff[j], (size_t)p.sendCount[j], ncclInt, j, d.comm, ctx->stream));
    }
    MOF_NCCL(ncclGroupEnd());
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

}  // namespace

static void p2p_release(mof_ctx* ctx, bool together);

bool dist_active(const mof_ctx* ctx) { return ctx->dist && ctx->dist->comm && ctx->dist->meshReady; }
int dist_world(const mof_ctx* ctx) { return ctx->dist ? ctx->dist->world : 1; }

void dist_range(const mof_ctx* ctx, int kind, int* s0, int* s1, int* r0, int* r1) {
    const DistState& d = *ctx->dist;
    *s0 = kind == 0 ? d.sliceStart[d.rank] : 0, *s1 = kind == 0 ? d.sliceStart[d.rank + 1] : 0;
    *r0 = d.part[kind].rowStart[d.rank], *r1 = d.part[kind].rowStart[d.rank + 1];
}

int dist_unique_id(unsigned char* id128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return MOF_E_CUDA;
    memcpy(id128, &id, sizeof(id));
    return MOF_OK;
}

int dist_init(mof_ctx* ctx, int world, int rank, const unsigned char* id128) {
    if (world < 1 || rank < 0 || rank >= world || !id128) return fail(ctx, MOF_E_INVALID, "mof_dist_init: bad world / rank / id");
    dist_destroy(ctx);
    ctx->dist = new DistState();
    DistState& d = *ctx->dist;
    d.world = world, d.rank = rank;
    d.part[0].width = 1, d.part[1].width = 6;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    MOF_NCCL(ncclCommInitRank(&d.comm, world, id, rank));
    const char* tr = getenv("MOF_DIST_TRACE");
    d.trace = tr && *tr && *tr != '0';
    if (d.trace) cudaEventCreate(&d.t0), cudaEventCreate(&d.t1);
    return MOF_OK;
}

void dist_destroy(mof_ctx* ctx) {
    if (!ctx->dist) return;
    DistState& d = *ctx->dist;
    if (d.trace && d.rank == 0)
        fprintf(stderr, "[dist trace] halo exchanges %lld: %.1f ms (%.1f us each, %.0f bytes each way on average); all-reduces %lld: %.1f ms (%.1f us each); gathers %lld: %.1f ms (%.1f us each)\n",
                d.traceN[TR_HALO], d.traceMs[TR_HALO], 1e3 * d.traceMs[TR_HALO] / std::max(1ll, d.traceN[TR_HALO]),
                d.traceMs[TR_HALO_BYTES] / std::max(1ll, d.traceN[TR_HALO_BYTES]) / 2, d.traceN[TR_ALLREDUCE], d.traceMs[TR_ALLREDUCE],
                1e3 * d.traceMs[TR_ALLREDUCE] / std::max(1ll, d.traceN[TR_ALLREDUCE]), d.traceN[TR_ALLGATHER], d.traceMs[TR_ALLGATHER],
                1e3 * d.traceMs[TR_ALLGATHER] / std::max(1ll, d.traceN[TR_ALLGATHER]));
    if (d.t0) cudaEventDestroy(d.t0), cudaEventDestroy(d.t1);
    for (Partition& p : d.part) p.sendIdx.release(), p.recvIdx.release(), p.offs.release();
    dist_clear_partitions(ctx);
    d.sendBuf.release(), d.recvBuf.release();
    if (d.comm) cudaStreamSynchronize(ctx->stream);
    p2p_release(ctx, d.win.shared);  // (contexts of a communicator are destroyed together, like the communicator itself)
    if (d.comm) {
        cudaStreamSynchronize(ctx->stream);
        ncclCommDestroy(d.comm);
    }
    delete ctx->dist;
    ctx->dist = nullptr;
}

// Row blocks and halo lists of both systems (patterns only: once per mesh).
int dist_setup_mesh(mof_ctx* ctx) {
    if (!ctx->dist || !ctx->dist->comm) return MOF_OK;
    DistState& d = *ctx->dist;
    d.meshReady = false;
    d.win.on = false;  // until dist_p2p_setup has seen the new mesh's partitions
    dist_clear_partitions(ctx);
    const int N = d.world, E = ctx->E, V = ctx->V, S = ctx->wSlices;
    d.sliceStart.assign(N + 1, 0);
    d.part[0].rowStart.assign(N + 1, 0), d.part[1].rowStart.assign(N + 1, 0);
    for (int k = 0; k <= N; k++) {
        d.sliceStart[k] = (int)((long long)S * k / N);
        d.part[0].rowStart[k] = std::min(E, 32 * d.sliceStart[k]);
        d.part[1].rowStart[k] = (int)((long long)V * k / N);
    }
    d.part[0].rowStart[N] = E;
    for (Partition& p : d.part) {
        p.sendCount.assign(N, 0), p.sendOff.assign(N, 0), p.recvCount.assign(N, 0), p.recvOff.assign(N, 0);
        p.nSend = p.nRecv = 0;
    }
    if (N == 1) {
        d.meshReady = true;
        return MOF_OK;
    }
    PhaseTimer pt(ctx);
    const int n = std::max(E, V);
    MOF_CUDA(ctx->itmp0.reserve((size_t)n + 1));  // scratch of the mesh set-up, free again at this point
    MOF_CUDA(ctx->itmp1.reserve((size_t)n + 1));
    {
        const int s0 = d.sliceStart[d.rank], s1 = d.sliceStart[d.rank + 1], r0 = d.part[0].rowStart[d.rank], r1 = d.part[0].rowStart[d.rank + 1];
        MOF_CUDA(cudaMemsetAsync(ctx->itmp0.p, 0, sizeof(int) * ((size_t)E + 1), ctx->stream));
        if (s1 > s0) MOF_LAUNCH(k_mark_halo_sell, kSMs * 8, B, 0, ctx->wSliceBase.p, ctx->wCol.p, s0, s1, r0, r1, ctx->itmp0.p);
        MOF_TRY(build_lists(ctx, d.part[0], E));
    }
    pt.mark("  halo lists of the flow system");
    {
        const int r0 = d.part[1].rowStart[d.rank], r1 = d.part[1].rowStart[d.rank + 1];
        MOF_CUDA(cudaMemsetAsync(ctx->itmp0.p, 0, sizeof(int) * ((size_t)V + 1), ctx->stream));
        if (r1 > r0) MOF_LAUNCH(k_mark_halo_csr, blocks_for(r1 - r0, B), B, 0, ctx->sRowptr.p, ctx->sCol.p, r0, r1, ctx->itmp0.p);
        MOF_TRY(build_lists(ctx, d.part[1], V));
    }
    pt.mark("  halo lists of the smoothing systems");
    size_t most = 1;
    for (Partition& p : d.part) most = std::max(most, (size_t)std::max(p.nSend, p.nRecv) * p.width);
    MOF_CUDA(d.sendBuf.reserve(most));
    MOF_CUDA(d.recvBuf.reserve(most));
    ctx->stats.haloEntries = d.part[0].nRecv;
    d.meshReady = true;
    return MOF_OK;
}

// ---- the peer-memory window (see PeerWindow): after every partition of the mesh exists (the matrix patterns' and the multigrid levels').
// together: every rank of the communicator is making this same call now (common knowledge at the call site, never a local condition).
static void p2p_release(mof_ctx* ctx, bool together) {
    DistState& d = *ctx->dist;
    PeerWindow& w = d.win;
    for (int j = 0; j < (int)w.peer.size(); j++)
        if (j != d.rank && w.peer[j]) cudaIpcCloseMemHandle(w.peer[j]);
    w.peer.clear();
    if (together && d.comm && d.world > 1) {
        // An exported allocation must outlive its mappings in the peers: every rank gets here at the same point of the program (a new
        // mesh, a failed set-up, the end of the context), closes its mappings above, and only frees its own window once all have.
        ScopedBuf<int> token;
        if (token.alloc(1) == cudaSuccess && cudaMemsetAsync(token.p, 0, sizeof(int), ctx->stream) == cudaSuccess &&
            ncclAllReduce(token.p, token.p, 1, ncclInt, ncclSum, d.comm, ctx->stream) == ncclSuccess)
            cudaStreamSynchronize(ctx->stream);
    }
    if (w.base) cudaFree(w.base);
    w.base = nullptr, w.bytes = 0, w.on = false, w.shared = false;
    w.table.release(), w.seq.release();
}

int dist_p2p_setup(mof_ctx* ctx) {
    if (!ctx->dist || !ctx->dist->comm || ctx->dist->world == 1) return MOF_OK;
    DistState& d = *ctx->dist;
    const int N = d.world;
    // Opt-in (MOF_DIST_P2P=1): measured on 2 and 4 B200 it is as fast as the NCCL exchange and no faster, and a library call
    // that every deployment already trusts is the better default for something that buys nothing.
    const char* e = getenv("MOF_DIST_P2P");
    const bool wanted = e && *e == '1';
    // what the widest message of any partition needs (fp64 bytes), the same number on every rank
    std::vector<Partition*> parts = {&d.part[0], &d.part[1]};
    for (Partition* q : d.extra) parts.push_back(q);
    long long need = 0;
    for (Partition* q : parts)
        for (int j = 0; j < N; j++) need = std::max(need, (long long)std::max(q->recvCount[j], q->sendCount[j]) * q->width * 8);
    // the solvers' own partitions are also all-gathered (a rank's range of two fp32 vectors per message): room for those of up to 4 MB,
    // larger ones keep going through NCCL (dist_allgather_part_f32 checks)
    for (Partition* q : d.extra)
        for (int j = 0; j < N; j++) {
            const long long bytes = (long long)(q->rowStart[j + 1] - q->rowStart[j]) * q->width * 8;
            if (bytes <= (4ll << 20)) need = std::max(need, bytes);
        }
    // (all-gathers of ints, reduced on the host: the message sizes are bytes of one halo message and fit an int)
    ScopedBuf<int> agreeSend, agreeAll;
    MOF_CUDA(agreeSend.alloc(2));
    MOF_CUDA(agreeAll.alloc(2 * (size_t)N));
    std::vector<int> gathered(2 * (size_t)N);
    auto agree = [&](int a, int b, long long* maxA, long long* minB) -> int {
        const int mine2[2] = {a, b};
        MOF_CUDA(cudaMemcpyAsync(agreeSend.p, mine2, sizeof(mine2), cudaMemcpyHostToDevice, ctx->stream));
        MOF_NCCL(ncclAllGather(agreeSend.p, agreeAll.p, 2, ncclInt, d.comm, ctx->stream));
        MOF_CUDA(read_back(ctx, gathered.data(), agreeAll.p, gathered.size()));
        *maxA = gathered[0], *minB = gathered[1];
        for (int j = 1; j < N; j++) *maxA = std::max<long long>(*maxA, gathered[2 * j]), *minB = std::min<long long>(*minB, gathered[2 * j + 1]);
        return MOF_OK;
    };
    long long all[2] = {0, 0};
    MOF_TRY(agree((int)std::min<long long>(need, 2000000000ll), wanted ? 1 : 0, &all[0], &all[1]));
    if (!all[1]) {  // switched off on some rank: NCCL everywhere
        p2p_release(ctx, d.win.shared);
        return MOF_OK;
    }
    const size_t cap = ((size_t)std::max(all[0], 4096ll) + 255) & ~(size_t)255;
    PeerWindow& w = d.win;
    long long ok = 1;
    if (!w.base || w.haloCap < cap) {
        // (re)create: every rank gets here together (the all-gathers above; capacities are common knowledge), and the stream is idle after the read-back
        p2p_release(ctx, d.win.shared);
        w.flagBytes = ((size_t)N * 2 * sizeof(unsigned long long) + 255) & ~(size_t)255;
        w.haloCap = cap;
        w.bytes = w.flagBytes + (size_t)N * 2 * cap;
        cudaIpcMemHandle_t handle;
        ScopedBuf<unsigned char> hsend, hall;
        std::vector<unsigned char> handles((size_t)N * sizeof(handle));
        if (cudaMalloc((void**)&w.base, w.bytes) != cudaSuccess || cudaMemset(w.base, 0, w.bytes) != cudaSuccess || cudaIpcGetMemHandle(&handle, w.base) != cudaSuccess) ok = 0;
        cudaGetLastError();
        static_assert(sizeof(handle) % sizeof(int) == 0, "the handles travel as ints");
        MOF_CUDA(hsend.alloc(sizeof(handle)));
        MOF_CUDA(hall.alloc((size_t)N * sizeof(handle)));
        if (!ok) memset(&handle, 0, sizeof(handle));
        MOF_CUDA(cudaMemcpyAsync(hsend.p, &handle, sizeof(handle), cudaMemcpyHostToDevice, ctx->stream));
        MOF_NCCL(ncclAllGather(hsend.p, hall.p, sizeof(handle) / sizeof(int), ncclInt, d.comm, ctx->stream));
        MOF_CUDA(read_back(ctx, handles.data(), hall.p, handles.size()));
        w.peer.assign(N, nullptr);
        for (int j = 0; j < N && ok; j++) {
            if (j == d.rank) { w.peer[j] = w.base; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, handles.data() + (size_t)j * sizeof(h), sizeof(h));
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0, cudaGetLastError();
            w.peer[j] = (unsigned char*)ptr;
        }
        if (ok) {
            MOF_CUDA(w.table.alloc((size_t)N));
            MOF_CUDA(w.seq.alloc((size_t)2 * N + 4));
            MOF_CUDA(cudaMemcpyAsync(w.table.p, w.peer.data(), sizeof(unsigned char*) * N, cudaMemcpyHostToDevice, ctx->stream));
            MOF_CUDA(cudaMemsetAsync(w.seq.p, 0, sizeof(unsigned long long) * (2 * N + 4), ctx->stream));
            MOF_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    // every rank mapped every window, or nobody uses them
    MOF_TRY(agree(0, (int)ok, &all[0], &all[1]));
    if (!all[1]) {
        p2p_release(ctx, true);
        if (getenv("MOF_MG_VERBOSE") && d.rank == 0) fprintf(stderr, "[dist] peer windows could not be mapped: halo exchanges go through NCCL\n");
        return MOF_OK;
    }
    for (Partition* q : parts) {
        std::vector<int> offs(3 * (size_t)(N + 1));
        for (int j = 0; j < N; j++) offs[j] = q->sendOff[j], offs[N + 1 + j] = q->recvOff[j];
        offs[N] = q->nSend, offs[2 * N + 1] = q->nRecv;
        for (int j = 0; j <= N; j++) offs[2 * (N + 1) + j] = q->rowStart[j];
        MOF_CUDA(q->offs.alloc(offs.size()));
        MOF_CUDA(cudaMemcpyAsync(q->offs.p, offs.data(), sizeof(int) * offs.size(), cudaMemcpyHostToDevice, ctx->stream));
        MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    w.on = w.shared = true;
    if (getenv("MOF_MG_VERBOSE") && d.rank == 0)
        fprintf(stderr, "[dist] halo exchanges through peer memory: %d windows of %.1f MB (%zu bytes per message slot)\n", N, w.bytes / 1048576., w.haloCap);
    return MOF_OK;
}
// A peer that never arrived (k_halo_wait gave up after ~10 s): the solve's result is garbage and says so.
int dist_p2p_check(mof_ctx* ctx) {
    if (!ctx->dist || !ctx->dist->win.on) return MOF_OK;
    DistState& d = *ctx->dist;
    unsigned long long flag = 0;
    MOF_CUDA(read_back(ctx, &flag, d.win.seq.p + 2 * d.world + 2));
    if (flag) return fail(ctx, MOF_E_CUDA, "[ERROR] partitioned mesh: a peer's halo values never arrived (peer-memory exchange timed out)");
    return MOF_OK;
}

int dist_halo_f64(mof_ctx* ctx, int kind, double* vec) { return halo_exchange<double>(ctx, ctx->dist->part[kind], vec, ncclDouble); }
int dist_halo_f32(mof_ctx* ctx, int kind, float* vec) { return halo_exchange<float>(ctx, ctx->dist->part[kind], vec, ncclFloat); }

// ---- partitions made by the solvers: any index set dealt to the ranks in contiguous ranges (multigrid levels by cell ranges, the
// fine rows a rank's aggregates reach into). The caller flags, in ctx->itmp0[0 .. n), the indices outside its own range that it
// reads; the lists are exchanged like those of the two matrix patterns.
int dist_add_partition(mof_ctx* ctx, int width, const int* rangeStart, int n, int* idOut) {
    DistState& d = *ctx->dist;
    Partition* p = new Partition();
    p->width = width;
    p->rowStart.assign(rangeStart, rangeStart + d.world + 1);
    d.extra.push_back(p);
    *idOut = (int)d.extra.size() - 1;
    if (d.world > 1) MOF_TRY(build_lists(ctx, *p, n));
    else p->sendCount.assign(1, 0), p->sendOff.assign(1, 0), p->recvCount.assign(1, 0), p->recvOff.assign(1, 0);
    const size_t most = (size_t)std::max(p->nSend, p->nRecv) * width;  // (in doubles: twice what the fp32 exchanges need)
    if (most > d.sendBuf.n) {
        MOF_CUDA(d.sendBuf.reserve(most));
        MOF_CUDA(d.recvBuf.reserve(most));
    }
    return MOF_OK;
}
void dist_clear_partitions(mof_ctx* ctx) {
    if (!ctx->dist) return;
    for (Partition* p : ctx->dist->extra) {
        p->sendIdx.release(), p->recvIdx.release(), p->offs.release();
        delete p;
    }
    ctx->dist->extra.clear();
}
long long dist_partition_halo(const mof_ctx* ctx, int id) { return ctx->dist->extra[id]->nRecv; }
int dist_halo_part_f32(mof_ctx* ctx, int id, float* vec) { return halo_exchange<float>(ctx, *ctx->dist->extra[id], vec, ncclFloat); }
// Every rank's own range of each of `count` vectors to all ranks, one group.
int dist_allgather_part_f32(mof_ctx* ctx, int id, float* const* vecs, int count) {
    DistState& d = *ctx->dist;
    if (d.world == 1) return MOF_OK;
    const Partition& p = *d.extra[id];
    TraceScope ts(ctx, TR_ALLGATHER);
    if (d.win.on && count <= 2 && p.offs.p) {
        long long widest = 0;
        for (int k = 0; k < d.world; k++) widest = std::max(widest, (long long)(p.rowStart[k + 1] - p.rowStart[k]) * p.width * (long long)sizeof(float) * count);
        if ((size_t)widest <= d.win.haloCap) {  // (the same decision on every rank: the ranges are common knowledge)
            const PeerWindow& win = d.win;
            GatherVecs vs;
            vs.count = count, vs.v[0] = vecs[0], vs.v[1] = count > 1 ? vecs[1] : nullptr;
            const int* rowStart = p.offs.p + 2 * (d.world + 1);
            const long long own = (long long)(p.rowStart[d.rank + 1] - p.rowStart[d.rank]) * p.width * count;
            const long long others = ((long long)p.rowStart[d.world] - (p.rowStart[d.rank + 1] - p.rowStart[d.rank])) * p.width * count;
            MOF_LAUNCH(k_gather_put, std::max(1, std::min(kSMs * 2, blocks_for(own, B))), B, 0, vs, rowStart, d.world, d.rank, p.width, (unsigned char* const*)win.table.p,
                       win.flagBytes, win.haloCap, win.seq.p);
            MOF_LAUNCH(k_gather_wait, std::max(1, std::min(kSMs * 2, blocks_for(others, B))), B, 0, vs, rowStart, d.world, d.rank, p.width, (const unsigned char*)win.base,
                       win.flagBytes, win.haloCap, win.seq.p);
            return MOF_OK;
        }
    }
    MOF_NCCL(ncclGroupStart());
    for (int v = 0; v < count; v++)
        for (int k = 0; k < d.world; k++) {
            const size_t n = (size_t)(p.rowStart[k + 1] - p.rowStart[k]) * p.width;
            float* at = vecs[v] + (size_t)p.rowStart[k] * p.width;
            if (n > 0) MOF_NCCL(ncclBroadcast(at, at, n, ncclFloat, k, d.comm, ctx->stream));
        }
    MOF_NCCL(ncclGroupEnd());
    return MOF_OK;
}
int dist_rank(const mof_ctx* ctx) { return ctx->dist ? ctx->dist->rank : 0; }
void dist_row_starts(const mof_ctx* ctx, int kind, int* out) {
    const std::vector<int>& r = ctx->dist->part[kind].rowStart;
    std::copy(r.begin(), r.end(), out);
}

int dist_allreduce_f64(mof_ctx* ctx, double* v, int count) {
    DistState& d = *ctx->dist;
    if (d.world == 1) return MOF_OK;
    TraceScope ts(ctx, TR_ALLREDUCE);
    MOF_NCCL(ncclAllReduce(v, v, (size_t)count, ncclDouble, ncclSum, d.comm, ctx->stream));
    return MOF_OK;
}
int dist_allreduce_f32(mof_ctx* ctx, float* v, int count) {
    DistState& d = *ctx->dist;
    if (d.world == 1) return MOF_OK;
    TraceScope ts(ctx, TR_ALLREDUCE);
    MOF_NCCL(ncclAllReduce(v, v, (size_t)count, ncclFloat, ncclSum, d.comm, ctx->stream));
    return MOF_OK;
}

// Every rank's own rows of `vec` to all ranks (the blocks have different lengths: one broadcast per block, grouped).
int dist_allgather_rows(mof_ctx* ctx, int kind, double* vec) {
    DistState& d = *ctx->dist;
    if (d.world == 1) return MOF_OK;
    const Partition& p = d.part[kind];
    MOF_NCCL(ncclGroupStart());
    for (int k = 0; k < d.world; k++) {
        const size_t n = (size_t)(p.rowStart[k + 1] - p.rowStart[k]) * p.width;
        double* at = vec + (size_t)p.rowStart[k] * p.width;
        if (n > 0) MOF_NCCL(ncclBroadcast(at, at, n, ncclDouble, k, d.comm, ctx->stream));
    }
    MOF_NCCL(ncclGroupEnd());
    return MOF_OK;
}

}  // namespace mof
