// Numbering of the caller's mesh (mof_set_mesh). Every kernel of the path gathers through vertex / triangle / edge indices
// (the SpMV's input vector, the aggregates of the multigrid levels, the walks' edge transforms), so what the caches do for
// them depends on how local the caller's numbering is — which a PLY file does not promise (the reference does not care: its
// factorisation applies a fill-reducing ordering of its own, LinearSolvers.h:277). Measured at 1 048 578 vertices on a B200
// (profiles/r2t): numbered along a space-filling curve an UpdateFlow iteration takes 46 ms, in the order a subdivision creates
// the vertices 71 ms, in random order 83 ms (SpMV 80 -> 92 us, walk 1.2 -> 3.9 ms).
//
// So a mesh whose numbering is not local is renumbered on the way in: vertices along the Morton curve of their positions
// (21 bits per axis of the bounding box), triangles along the Morton curve of their centroids, both by a stable radix sort
// of (code, index) pairs. Everything downstream runs in the new numbering; per-vertex inputs (the signals) are gathered
// and per-vertex / per-triangle outputs (advected colours, the flow) scattered back at the C ABI, so the caller never sees it.
// A triangle keeps its corner order, hence its chart: per-triangle 2-vectors need no change, only a new place.
// "Not local" = the mean index span of a triangle's corners, or the mean jump between the first corners of consecutive
// triangles, exceeds V/16 (a curve-ordered mesh: < V/60 at 16 k vertices, falling with size; creation order or random: V/2).
// Small meshes (< 65 536 vertices) live in L2 whatever their numbering and are left alone.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "mof_internal.cuh"

#ifndef MOF_HOST_EMULATION
#include <cub/device/device_radix_sort.cuh>
#endif

namespace mof {

namespace {

constexpr int B = 256;

__global__ void k_bbox_partial(const double* __restrict__ pos, int V, double* __restrict__ partial) {
    __shared__ double sh[6][B];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x)
        for (int k = 0; k < 3; k++) lo[k] = fmin(lo[k], pos[3 * v + k]), hi[k] = fmax(hi[k], pos[3 * v + k]);
    for (int k = 0; k < 3; k++) sh[k][threadIdx.x] = lo[k], sh[3 + k][threadIdx.x] = hi[k];
    __syncthreads();
    for (int s = B / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
            for (int k = 0; k < 3; k++) {
                sh[k][threadIdx.x] = fmin(sh[k][threadIdx.x], sh[k][threadIdx.x + s]);
                sh[3 + k][threadIdx.x] = fmax(sh[3 + k][threadIdx.x], sh[3 + k][threadIdx.x + s]);
            }
        __syncthreads();
    }
    if (threadIdx.x < 6) partial[6 * blockIdx.x + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void k_bbox_final(const double* __restrict__ partial, int np, double* __restrict__ box) {
    if (threadIdx.x >= 6) return;
    const int k = threadIdx.x;
    double v = partial[k];
    for (int i = 1; i < np; i++) v = k < 3 ? fmin(v, partial[6 * i + k]) : fmax(v, partial[6 * i + k]);
    box[k] = v;
}

__host__ __device__ __forceinline__ unsigned long long spread21(unsigned long long x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__device__ __forceinline__ unsigned long long morton63(const double* box, double x, double y, double z) {
    const double p[3] = {x, y, z};
    unsigned long long q[3];
    for (int k = 0; k < 3; k++) {
        const double w = box[3 + k] - box[k];
        double u = w > 0 ? (p[k] - box[k]) / w : 0.;
        u = fmin(1., fmax(0., u));
        q[k] = (unsigned long long)(u * 2097151.);
    }
    return spread21(q[0]) | spread21(q[1]) << 1 | spread21(q[2]) << 2;
}
__global__ void k_vertex_codes(const double* __restrict__ pos, const double* __restrict__ box, int V, unsigned long long* __restrict__ code, int* __restrict__ idx) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    code[v] = morton63(box, pos[3 * v], pos[3 * v + 1], pos[3 * v + 2]);
    idx[v] = v;
}
__global__ void k_triangle_codes(const double* __restrict__ pos, const int* __restrict__ tri, const double* __restrict__ box, int T, unsigned long long* __restrict__ code,
                                 int* __restrict__ idx) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double c[3] = {0, 0, 0};
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) c[k] += pos[3 * (size_t)tri[3 * t + j] + k];
    code[t] = morton63(box, c[0] / 3, c[1] / 3, c[2] / 3);
    idx[t] = t;
}
// Locality of the caller's numbering, per triangle: the index span of its corners and the jump from the previous triangle's first corner.
__global__ void k_locality(const int* __restrict__ tri, int T, int V, double* __restrict__ span, double* __restrict__ jump) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
    // (indices outside [0, V) are reported by the operator assembly; here they must only not fault)
    span[t] = (double)(max(a, max(b, c)) - min(a, min(b, c)));
    jump[t] = t ? fabs((double)a - (double)tri[3 * t - 3]) : 0.;
}
__global__ void k_invert_order(const int* __restrict__ order, int n, int* __restrict__ rank) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rank[order[i]] = i;
}
__global__ void k_gather_positions(const double* __restrict__ in, const int* __restrict__ order, int V, double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3ll * V) return;
    long long v = i / 3, k = i - 3 * v;
    out[i] = in[3 * (size_t)order[v] + k];
}
__global__ void k_gather_triangles(const int* __restrict__ in, const int* __restrict__ tOrder, const int* __restrict__ vRank, int T, int V, int* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3ll * T) return;
    long long t = i / 3, j = i - 3 * t;
    const int v = in[3 * (size_t)tOrder[t] + j];
    out[i] = v >= 0 && v < V ? vRank[v] : v;  // (a bad index stays bad: build_mesh_operators reports it)
}
// out[(dst of i)][k] = in[(src of i)][k], rows of `width` doubles: gather (order = new -> old, rows of `out` in the new numbering) ...
__global__ void k_gather_rows(const double* __restrict__ in, const int* __restrict__ order, long long n, int width, double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * width) return;
    long long r = i / width, k = i - r * width;
    out[i] = in[(size_t)order[r] * width + k];
}
// ... and scatter (rows of `in` in the new numbering, `out` in the caller's)
__global__ void k_scatter_rows(const double* __restrict__ in, const int* __restrict__ order, long long n, int width, double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * width) return;
    long long r = i / width, k = i - r * width;
    out[(size_t)order[r] * width + k] = in[i];
}

// Stable sort of (code, index) pairs by code; `idx` ends as the order (new -> old).
int sort_pairs(mof_ctx* ctx, unsigned long long* code, int* idx, int n) {
#ifdef MOF_HOST_EMULATION
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<std::pair<unsigned long long, int>> p((size_t)n);
    for (int i = 0; i < n; i++) p[i] = {code[i], idx[i]};
    std::stable_sort(p.begin(), p.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    for (int i = 0; i < n; i++) code[i] = p[i].first, idx[i] = p[i].second;
    return MOF_OK;
#else
    ScopedBuf<unsigned long long> code2;
    ScopedBuf<int> idx2;
    ScopedBuf<unsigned char> scratch;
    MOF_CUDA(code2.alloc((size_t)n));
    MOF_CUDA(idx2.alloc((size_t)n));
    size_t bytes = 0;
    MOF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, code, code2.p, idx, idx2.p, n, 0, 63, ctx->stream));
    MOF_CUDA(scratch.alloc(bytes));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(scratch.p, bytes, code, code2.p, idx, idx2.p, n, 0, 63, ctx->stream);  // (LSD radix sort: stable)
    if (e == cudaSuccess) e = cudaMemcpyAsync(idx, idx2.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream);
    MOF_CUDA(e);
    return MOF_OK;
#endif
}

}  // namespace

// Called by mof_set_mesh with the caller's mesh in ctx->pos / ctx->tri: decides, and if the numbering is not local replaces both by
// the renumbered mesh and keeps the two orders (new -> old). mode: 0 never, 1 always, anything else by the measure above.
int reorder_mesh(mof_ctx* ctx, int mode) {
    const int V = ctx->V, T = ctx->T;
    ctx->reordered = false;
    if (mode == 0) return MOF_OK;
    if (mode != 1) {
        const char* e = getenv("MOF_REORDER_MIN_VERTICES");  // (tests: the decision itself at sizes the CPU tier runs)
        if (V < (e && *e ? atoi(e) : 65536)) return MOF_OK;
        MOF_CUDA(ctx->scalars.alloc(SC_COUNT));  // (the operator assembly allocates it too; this may be the context's first mesh)
        MOF_CUDA(ctx->dtmp0.reserve(2ull * T));
        MOF_LAUNCH(k_locality, blocks_for(T, B), B, 0, ctx->tri.p, T, V, ctx->dtmp0.p, ctx->dtmp0.p + T);
        MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, T, ctx->scalars.p + SC_TMP));
        MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p + T, T, ctx->scalars.p + SC_TMP + 1));
        double sums[2] = {0, 0};
        MOF_CUDA(read_back(ctx, sums, ctx->scalars.p + SC_TMP, 2));
        const double limit = (double)V / 16.;
        if (sums[0] / T <= limit && sums[1] / T <= limit) return MOF_OK;
    }
    ScopedBuf<double> box, partial, pos2;
    ScopedBuf<unsigned long long> code;
    ScopedBuf<int> tri2;
    const int np = std::min(1024, blocks_for(V, B));
    MOF_CUDA(box.alloc(6));
    MOF_CUDA(partial.alloc(6ull * np));
    MOF_CUDA(code.alloc((size_t)std::max(V, T)));
    MOF_CUDA(ctx->vOrder.alloc((size_t)V));
    MOF_CUDA(ctx->vRank.alloc((size_t)V));
    MOF_CUDA(ctx->tOrder.alloc((size_t)T));
    MOF_LAUNCH(k_bbox_partial, np, B, 0, ctx->pos.p, V, partial.p);
    MOF_LAUNCH(k_bbox_final, 1, 32, 0, partial.p, np, box.p);
    MOF_LAUNCH(k_vertex_codes, blocks_for(V, B), B, 0, ctx->pos.p, box.p, V, code.p, ctx->vOrder.p);
    MOF_TRY(sort_pairs(ctx, code.p, ctx->vOrder.p, V));
    MOF_LAUNCH(k_invert_order, blocks_for(V, B), B, 0, ctx->vOrder.p, V, ctx->vRank.p);
    // triangles by their centroids (positions still in the caller's numbering, like the indices in ctx->tri)
    MOF_LAUNCH(k_triangle_codes, blocks_for(T, B), B, 0, ctx->pos.p, ctx->tri.p, box.p, T, code.p, ctx->tOrder.p);
    MOF_TRY(sort_pairs(ctx, code.p, ctx->tOrder.p, T));
    MOF_CUDA(pos2.alloc(3ull * V));
    MOF_CUDA(tri2.alloc(3ull * T));
    MOF_LAUNCH(k_gather_positions, blocks_for(3ll * V, B), B, 0, ctx->pos.p, ctx->vOrder.p, V, pos2.p);
    MOF_LAUNCH(k_gather_triangles, blocks_for(3ll * T, B), B, 0, ctx->tri.p, ctx->tOrder.p, ctx->vRank.p, T, V, tri2.p);
    MOF_CUDA(cudaMemcpyAsync(ctx->pos.p, pos2.p, sizeof(double) * 3 * V, cudaMemcpyDeviceToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->tri.p, tri2.p, sizeof(int) * 3 * T, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->reordered = true;
    return MOF_OK;
}

// Rows of `width` doubles between the caller's numbering and the library's. kind 0: vertices, 1: triangles.
int reorder_gather(mof_ctx* ctx, int kind, const double* callerRows, int width, double* libraryRows) {
    const int* order = kind == 0 ? ctx->vOrder.p : ctx->tOrder.p;
    const long long n = kind == 0 ? ctx->V : ctx->T;
    MOF_LAUNCH(k_gather_rows, blocks_for(n * width, B), B, 0, callerRows, order, n, width, libraryRows);
    return MOF_OK;
}
int reorder_scatter(mof_ctx* ctx, int kind, const double* libraryRows, int width, double* callerRows) {
    const int* order = kind == 0 ? ctx->vOrder.p : ctx->tOrder.p;
    const long long n = kind == 0 ? ctx->V : ctx->T;
    MOF_LAUNCH(k_scatter_rows, blocks_for(n * width, B), B, 0, libraryRows, order, n, width, callerRows);
    return MOF_OK;
}

}  // namespace mof
