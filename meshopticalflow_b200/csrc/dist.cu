// One mesh over several GPUs (BASELINE.json configs[4]): the rows of the flow system are split into
// contiguous blocks of 32-row slices, one block per rank (one process per GPU). Everything else — mesh operators,
// signals, walks, the coarse multigrid levels — is replicated: every rank makes the same calls with the same inputs
// and holds full-length vectors, of which it computes its own rows. What crosses NVLink, through NCCL on the
// context's stream:
//   * halo exchange before every fine-level SpMV: the entries of the input vector that a rank's rows reference outside
//     its block (index lists built once per mesh from the matrix pattern, grouped ncclSend/ncclRecv of packed values),
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

struct DistState {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    bool meshReady = false;
    std::vector<int> sliceStart;            // world + 1
    std::vector<int> rowStart;              // world + 1
    std::vector<int> sendCount, sendOff, recvCount, recvOff;  // per peer, in entries
    int nSend = 0, nRecv = 0;
    DBuf<int> sendIdx, recvIdx;
    DBuf<double> sendBuf, recvBuf;          // sized for fp64 entries, reused for fp32
};

namespace {

#define MOF_NCCL(call)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) return fail(ctx, MOF_E_CUDA, std::string(#call) + ": " + ncclGetErrorString(r__)); \
    } while (0)

constexpr int B = 256;

// flags[c] = 1 for every column outside [r0, r1) referenced by the slices [s0, s1)
__global__ void k_mark_halo(const int* __restrict__ sliceBase, const int* __restrict__ col, int s0, int s1, int r0, int r1, int* __restrict__ flags) {
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
__global__ void k_compact(const int* __restrict__ flags, const int* __restrict__ pos, int n, int* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[pos[i]] = i;
}
template <class T>
__global__ void k_pack(const T* __restrict__ vec, const int* __restrict__ idx, int n, T* __restrict__ buf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = vec[idx[i]];
}
template <class T>
__global__ void k_unpack(const T* __restrict__ buf, const int* __restrict__ idx, int n, T* __restrict__ vec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vec[idx[i]] = buf[i];
}

template <class T>
int halo_exchange(mof_ctx* ctx, T* vec, ncclDataType_t type) {
    DistState& d = *ctx->dist;
    if (d.world == 1) return MOF_OK;
    T* sb = (T*)d.sendBuf.p;
    T* rb = (T*)d.recvBuf.p;
    if (d.nSend) MOF_LAUNCH(k_pack<T>, blocks_for(d.nSend, B), B, 0, vec, d.sendIdx.p, d.nSend, sb);
    MOF_NCCL(ncclGroupStart());
    for (int j = 0; j < d.world; j++) {
        if (j == d.rank) continue;
        if (d.sendCount[j]) MOF_NCCL(ncclSend(sb + d.sendOff[j], (size_t)d.sendCount[j], type, j, d.comm, ctx->stream));
        if (d.recvCount[j]) MOF_NCCL(ncclRecv(rb + d.recvOff[j], (size_t)d.recvCount[j], type, j, d.comm, ctx->stream));
    }
    MOF_NCCL(ncclGroupEnd());
    if (d.nRecv) MOF_LAUNCH(k_unpack<T>, blocks_for(d.nRecv, B), B, 0, rb, d.recvIdx.p, d.nRecv, vec);
    return MOF_OK;
}

}  // namespace

bool dist_active(const mof_ctx* ctx) { return ctx->dist && ctx->dist->comm && ctx->dist->meshReady; }
int dist_world(const mof_ctx* ctx) { return ctx->dist ? ctx->dist->world : 1; }

void dist_range(const mof_ctx* ctx, int* s0, int* s1, int* r0, int* r1) {
    const DistState& d = *ctx->dist;
    *s0 = d.sliceStart[d.rank], *s1 = d.sliceStart[d.rank + 1];
    *r0 = d.rowStart[d.rank], *r1 = d.rowStart[d.rank + 1];
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
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    MOF_NCCL(ncclCommInitRank(&d.comm, world, id, rank));
    return MOF_OK;
}

void dist_destroy(mof_ctx* ctx) {
    if (!ctx->dist) return;
    DistState& d = *ctx->dist;
    d.sendIdx.release(), d.recvIdx.release(), d.sendBuf.release(), d.recvBuf.release();
    if (d.comm) {
        cudaStreamSynchronize(ctx->stream);
        ncclCommDestroy(d.comm);
    }
    delete ctx->dist;
    ctx->dist = nullptr;
}

// Row blocks and halo lists of the flow matrix (pattern only: once per mesh).
int dist_setup_mesh(mof_ctx* ctx) {
    if (!ctx->dist || !ctx->dist->comm) return MOF_OK;
    DistState& d = *ctx->dist;
    d.meshReady = false;
    const int N = d.world, E = ctx->E, S = ctx->wSlices;
    d.sliceStart.assign(N + 1, 0), d.rowStart.assign(N + 1, 0);
    for (int k = 0; k <= N; k++) {
        d.sliceStart[k] = (int)((long long)S * k / N);
        d.rowStart[k] = std::min(E, 32 * d.sliceStart[k]);
    }
    d.rowStart[N] = E;
    d.sendCount.assign(N, 0), d.sendOff.assign(N, 0), d.recvCount.assign(N, 0), d.recvOff.assign(N, 0);
    d.nSend = d.nRecv = 0;
    if (N == 1) {
        d.meshReady = true;
        return MOF_OK;
    }
    const int s0 = d.sliceStart[d.rank], s1 = d.sliceStart[d.rank + 1], r0 = d.rowStart[d.rank], r1 = d.rowStart[d.rank + 1];
    PhaseTimer pt(ctx);
    // columns my rows reference outside my block, ascending (so grouped by owner)
    DBuf<int>& flags = ctx->itmp0;  // scratch of the mesh set-up, free again at this point
    DBuf<int>& pos = ctx->itmp1;
    MOF_CUDA(flags.reserve((size_t)E + 1));
    MOF_CUDA(pos.reserve((size_t)E + 1));
    MOF_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int) * ((size_t)E + 1), ctx->stream));
    if (s1 > s0) MOF_LAUNCH(k_mark_halo, kSMs * 8, B, 0, ctx->wSliceBase.p, ctx->wCol.p, s0, s1, r0, r1, flags.p);
    MOF_TRY(exclusive_scan_int(ctx, flags.p, pos.p, E + 1, nullptr));
    int H = 0;
    MOF_CUDA(read_back(ctx, &H, pos.p + E));
    d.nRecv = H;
    MOF_CUDA(d.recvIdx.alloc((size_t)std::max(H, 1)));
    if (H) MOF_LAUNCH(k_compact, blocks_for(E, B), B, 0, flags.p, pos.p, E, d.recvIdx.p);
    std::vector<int> hIdx((size_t)H);
    if (H) MOF_CUDA(read_back(ctx, hIdx.data(), d.recvIdx.p, (size_t)H));
    pt.mark("  halo: columns outside my block");
    for (int k = 0, i = 0; k < N; k++) {
        d.recvOff[k] = i;
        while (i < H && hIdx[i] < d.rowStart[k + 1]) i++;
        d.recvCount[k] = i - d.recvOff[k];
    }
    // who needs how much from whom: row k of the matrix = recvCount of rank k
    DBuf<int> counts, matrix;
    MOF_CUDA(counts.alloc(N));
    MOF_CUDA(matrix.alloc((size_t)N * N));
    MOF_CUDA(cudaMemcpyAsync(counts.p, d.recvCount.data(), sizeof(int) * N, cudaMemcpyHostToDevice, ctx->stream));
    MOF_NCCL(ncclAllGather(counts.p, matrix.p, N, ncclInt, d.comm, ctx->stream));
    std::vector<int> hm((size_t)N * N);
    MOF_CUDA(read_back(ctx, hm.data(), matrix.p, (size_t)N * N));
    counts.release(), matrix.release();
    pt.mark("  halo: counts all-gather");
    for (int j = 0, off = 0; j < N; j++) {
        d.sendCount[j] = j == d.rank ? 0 : hm[(size_t)j * N + d.rank];
        d.sendOff[j] = off;
        off += d.sendCount[j];
        d.nSend = off;
    }
    MOF_CUDA(d.sendIdx.alloc((size_t)std::max(d.nSend, 1)));
    MOF_NCCL(ncclGroupStart());
    for (int j = 0; j < N; j++) {
        if (j == d.rank) continue;
        if (d.recvCount[j]) MOF_NCCL(ncclSend(d.recvIdx.p + d.recvOff[j], (size_t)d.recvCount[j], ncclInt, j, d.comm, ctx->stream));
        if (d.sendCount[j]) MOF_NCCL(ncclRecv(d.sendIdx.p + d.sendOff[j], (size_t)d.sendCount[j], ncclInt, j, d.comm, ctx->stream));
    }
    MOF_NCCL(ncclGroupEnd());
    MOF_CUDA(d.sendBuf.alloc((size_t)std::max(d.nSend, 1)));
    MOF_CUDA(d.recvBuf.alloc((size_t)std::max(d.nRecv, 1)));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    pt.mark("  halo: index lists exchange");
    ctx->stats.haloEntries = d.nRecv;
    d.meshReady = true;
    return MOF_OK;
}

int dist_halo_f64(mof_ctx* ctx, double* vec) { return halo_exchange<double>(ctx, vec, ncclDouble); }
int dist_halo_f32(mof_ctx* ctx, float* vec) { return halo_exchange<float>(ctx, vec, ncclFloat); }

int dist_allreduce_f64(mof_ctx* ctx, double* v, int count) {
    DistState& d = *ctx->dist;
    if (d.world == 1) return MOF_OK;
    MOF_NCCL(ncclAllReduce(v, v, (size_t)count, ncclDouble, ncclSum, d.comm, ctx->stream));
    return MOF_OK;
}
int dist_allreduce_f32(mof_ctx* ctx, float* v, int count) {
    DistState& d = *ctx->dist;
    if (d.world == 1) return MOF_OK;
    MOF_NCCL(ncclAllReduce(v, v, (size_t)count, ncclFloat, ncclSum, d.comm, ctx->stream));
    return MOF_OK;
}

// Every rank's own rows of `vec` to all ranks (the blocks have different lengths: one broadcast per block, grouped).
int dist_allgather_rows(mof_ctx* ctx, double* vec) {
    DistState& d = *ctx->dist;
    if (d.world == 1) return MOF_OK;
    MOF_NCCL(ncclGroupStart());
    for (int k = 0; k < d.world; k++) {
        const int n = d.rowStart[k + 1] - d.rowStart[k];
        if (n > 0) MOF_NCCL(ncclBroadcast(vec + d.rowStart[k], vec + d.rowStart[k], (size_t)n, ncclDouble, k, d.comm, ctx->stream));
    }
    MOF_NCCL(ncclGroupEnd());
    return MOF_OK;
}

}  // namespace mof
