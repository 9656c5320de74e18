// Internal declarations shared by the .cu files of libmof_b200.so. Not part of the C ABI.
#pragma once

// MOF_HOST_EMULATION (tests/host_emulation, CPU test tier only): the CUDA runtime calls and the launch syntax are
// replaced by host stand-ins so that whole .cu files can be compiled by g++ and run one "thread" at a time.
#ifdef MOF_HOST_EMULATION
#include <tuple>

#include "emul_cuda_runtime.h"
#else
#include <cuda_runtime.h>
#endif

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/mof_b200.h"

#include "mof_slots.h"

namespace mof {

// Device buffers come from CUDA's stream-ordered pool (cudaMallocAsync / cudaFreeAsync on the context's stream, release
// threshold raised in mof_create): re-running the setup for a new mesh or pair re-uses the pool's memory without
// device-wide synchronisation. The stream is the one of the C-ABI call in progress (set by StreamScope in mof_api.cu).
inline cudaStream_t& alloc_stream() {
    static thread_local cudaStream_t s = nullptr;
    return s;
}

// A buffer keeps its memory when it is re-sized to something that still fits (n = logical size, cap = allocated):
// aligning pair after pair re-runs the whole set-up with the same sizes, and handing hundreds of MB back to the pool
// only to ask for them again lets other allocations carve the block up, after which the pool has to grow — which
// costs seconds once NCCL has enabled peer access.
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    cudaError_t alloc(size_t count) {
        if (p && count <= cap && count > 0) {
            n = count;
            return cudaSuccess;
        }
        release();
        if (!count) return cudaSuccess;
        cudaError_t e = cudaMallocAsync((void**)&p, count * sizeof(T), alloc_stream());
        if (e == cudaSuccess) n = cap = count;
        else p = nullptr;
        return e;
    }
    // Scratch use: grow-only, the logical size follows the largest request.
    cudaError_t reserve(size_t count) { return (p && count <= n) ? cudaSuccess : alloc(std::max(count, n)); }
    void release() {
        if (p) cudaFreeAsync(p, alloc_stream());
        p = nullptr, n = cap = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};
// A scratch buffer of one function: handed back to the pool on every way out of the scope, error returns included.
template <class T>
struct ScopedBuf : DBuf<T> {
    ScopedBuf() = default;
    ScopedBuf(const ScopedBuf&) = delete;
    ScopedBuf& operator=(const ScopedBuf&) = delete;
    ~ScopedBuf() { this->release(); }
};

// Sliced storage of the E x E Whitney operators (SELL-32): rows are grouped by 32 (one warp), every
// group is padded to its longest row and stored entry-major, i.e. entry j of row r lives at
// sliceBase[r / 32] + 32 * j + r % 32. A warp that owns a slice reads 32 consecutive words per entry
// index with no staging and no barrier, and every thread has all of its row's loads independent and in
// flight at once. Padding entries carry value 0 and the row's own column.
#if defined(__CUDACC__) || defined(MOF_HOST_EMULATION)
__host__ __device__ __forceinline__ size_t sell_pos(const int* sliceBase, int row, int j) {
    return (size_t)sliceBase[row >> 5] + 32 * (size_t)j + (size_t)(row & 31);
}
#endif

struct Multigrid;  // multigrid.cu
struct DistState;  // dist.cu
struct VfState;    // vector_fields.cu
struct SmoothAhead;  // flow_kernels.cu

struct PcgWork {
    DBuf<double> r, d, q;       // [n * nrhs]
    DBuf<double> partial;       // block partials, 3 banks
    DBuf<double> result;        // [8]: iterations, relres, converged flag ...
    int gridBlocks = 0;
};

}  // namespace mof

struct mof_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int reorderMode = -1;  // mof_set_reorder: -1 by the locality of the caller's numbering, 0 never, 1 always (reorder.cu)
    bool reordered = false;  // the mesh in ctx->pos / ctx->tri is numbered along a Morton curve; vOrder / tOrder: new -> old, vRank: old -> new
    mof::DBuf<int> vOrder, vRank, tOrder;
    bool pdl = false;  // kernels of the solvers are launched with programmatic stream serialisation (MOF_LAUNCH_PDL; MOF_PDL=0 turns it off)
    std::string err;
    mof_params params;
    mof_stats stats;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double* pinned = nullptr;  // 64 doubles of page-locked host memory, allocated once (cudaHostAlloc / cudaFreeHost cost up to 0.4 s)

    int V = 0, T = 0, E = 0;
    long long nnzS = 0, nnzW = 0, wPadded = 0;
    int wSlices = 0;
    bool haveMesh = false, haveSignals = false, haveFlowSystem = false, haveTexture = false;
    int iterationsDone = 0;
    double curSmooth = 0, curVf = 0;

    // mesh
    mof::DBuf<double> pos, g, area, xlin, xcst;
    mof::DBuf<int> tri, opp;
    // scalar (V x V) operators, one pattern
    mof::DBuf<int> sRowptr, sCol, sHe;
    mof::DBuf<double> sMass, sStiff, sSys, sDinv;
    // ... and the system matrix once more in the sliced layout (SELL-32) for the solver's kernels: pattern per mesh, values per system
    mof::DBuf<int> sSliceBase, sColSell;
    mof::DBuf<double> sSysSell;
    long long sPadded = 0;
    // Whitney
    mof::DBuf<int> reduced, expanded, positive, wRowptr, wSliceBase, wCol;  // wCol/wS/wA: sliced layout (sell_pos), wPadded entries
    mof::DBuf<double> P, m0, m1, wS, wA, wDinv;
    // signals, (A rgb, B rgb) interleaved per vertex
    mof::DBuf<double> raw6, log6, sig6, smoothed6, rhs6, resampled6, tsample6, dataD, dataRhs;  // log6: --log transform of raw6 (comparison only)
    // flow unknowns
    mof::DBuf<double> coeffs, tfield, fb, fx;
    mof::DBuf<double> scalars;
    mof::PcgWork pcg;
    mof::Multigrid* mg = nullptr;   // multilevel preconditioner of the flow system
    mof::Multigrid* mgs = nullptr;  // ... and of the scalar smoothing systems
    mof::DistState* dist = nullptr; // one mesh over several GPUs (dist.cu); nullptr = single GPU
    mof::VfState* vf = nullptr;     // Conformal / Connection basis of the signals in place (vector_fields.cu); nullptr = Whitney
    mof::SmoothAhead* ahead = nullptr;  // the NEXT iteration's smoothing solve, running on a second stream under the flow solve
    // 6-channel blend (0 < dogWeight < 1, OpticalFlow.cpp:849-855): sig6 holds the DoG half (times w), these the raw half (times 1-w)
    bool blend = false;
    mof::DBuf<double> sigLo6, smoothedLo6, resampledLo6;
    // scratch
    mof::DBuf<int> itmp0, itmp1, itmp2, flags;
    mof::DBuf<unsigned long long> hashKeys;
    mof::DBuf<double> dtmp0, dtmp1, dtmp2;
    // texture path
    int texW = 0, texH = 0;
    mof::DBuf<int> srcT;
    mof::DBuf<double> srcP, triUV, texOut;
    mof::DBuf<unsigned char> tex[2];
    // texture-configuration preparation on the device (texprep_kernels.cu): the mesh being subdivided
    int subV = 0, subT = 0;
    mof::DBuf<float> subXyz;
    mof::DBuf<int> subTri;
    mof::DBuf<double> subUv;
};

namespace mof {

inline int fail(mof_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
inline int cuda_fail(mof_ctx* c, cudaError_t e, const char* what) {
    return fail(c, MOF_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define MOF_CUDA(call)                                                  \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return mof::cuda_fail(ctx, e__, #call); \
    } while (0)

#define MOF_TRY(call)             \
    do {                          \
        int rc__ = (call);        \
        if (rc__ != MOF_OK) return rc__; \
    } while (0)

inline int blocks_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }

// Host read of device values AFTER everything queued on the context's stream. (The context's stream is non-blocking:
// a plain cudaMemcpy on the legacy default stream would not be ordered against it.)
template <class T>
inline cudaError_t read_back(mof_ctx* ctx, T* host, const T* dev, size_t count = 1) {
    cudaError_t e = cudaMemcpyAsync(host, dev, sizeof(T) * count, cudaMemcpyDeviceToHost, ctx->stream);
    return e != cudaSuccess ? e : cudaStreamSynchronize(ctx->stream);
}

// MOF_VERBOSE_SETUP=1: wall-clock of the set-up phases on stderr (each mark synchronises the stream).
struct PhaseTimer {
    mof_ctx* ctx;
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(mof_ctx* c) : ctx(c) {
        const char* e = getenv("MOF_VERBOSE_SETUP");
        on = e && *e && *e != '0';
        if (on) cudaStreamSynchronize(ctx->stream), t0 = std::chrono::steady_clock::now();
    }
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(ctx->stream);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[setup] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Counts a kernel launch of ours (mof_stats.kernelLaunches) and checks the launch.
#ifdef MOF_HOST_EMULATION
#define MOF_LAUNCH(kernel, grid, block, smem, ...)                                                                                  \
    do {                                                                                                                            \
        auto args__ = std::make_tuple(__VA_ARGS__); /* by value, like a launch: a captured launch runs after this scope is gone */  \
        mof_emul::submit((grid), (block), [args__] { std::apply([](auto... a) { kernel(a...); }, args__); });                       \
        ctx->stats.kernelLaunches++;                                                                                                \
    } while (0)
#else
#define MOF_LAUNCH(kernel, grid, block, smem, ...)                                  \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);              \
        ctx->stats.kernelLaunches++;                                                \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) return mof::cuda_fail(ctx, e__, #kernel);           \
    } while (0)
#endif

// Programmatic dependent launch (griddepcontrol, sm_90+): a kernel launched through MOF_LAUNCH_PDL may have its CTAs scheduled while the
// previous kernel on the stream is still draining, which hides most of the ~2 us launch-to-launch gap of the dependent chain of small
// kernels a multigrid cycle is made of. The contract: such a kernel's FIRST statement is pdl_wait() — it returns when the preceding
// grids have completed and their writes are visible, so nothing before that point may touch memory (reads or writes); with every
// kernel of the chain doing so the ordering is that of plain stream order. Without the launch attribute the instruction is a no-op.
#ifdef MOF_HOST_EMULATION
inline void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif
#ifdef MOF_HOST_EMULATION
#define MOF_LAUNCH_PDL MOF_LAUNCH
#else
#define MOF_LAUNCH_PDL(kernel, grid, block, smem, ...)                                                  \
    do {                                                                                                \
        cudaError_t e__;                                                                                \
        if (ctx->pdl) {                                                                                 \
            cudaLaunchConfig_t cfg__ = {};                                                              \
            cfg__.gridDim = dim3((unsigned)(grid)), cfg__.blockDim = dim3((unsigned)(block));           \
            cfg__.dynamicSmemBytes = (smem), cfg__.stream = ctx->stream;                                \
            cudaLaunchAttribute attr__[1];                                                              \
            attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                          \
            attr__[0].val.programmaticStreamSerializationAllowed = 1;                                   \
            cfg__.attrs = attr__, cfg__.numAttrs = 1;                                                   \
            e__ = cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                                      \
        } else {                                                                                        \
            kernel<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);                              \
            e__ = cudaGetLastError();                                                                   \
        }                                                                                               \
        ctx->stats.kernelLaunches++;                                                                    \
        if (e__ != cudaSuccess) return mof::cuda_fail(ctx, e__, #kernel);                               \
    } while (0)
#endif

// spectrum.cu and what it borrows from the alignment path
int spectrum_lowest(mof_ctx* ctx, int count, double tol, int maxIterations, double* eigenvalues, double* fields, int* iterationsOut, double* residualOut);
int metric_mass_blocks(mof_ctx* ctx);                       // flow_kernels.cu: ctx->dataD = g_t area_t
int whitney_mass_operator(mof_ctx* ctx, double* wM);        // flow_kernels.cu
int whitney_triangle_field(mof_ctx* ctx, const double* coeffs, double* tfield);
int vf_apply_operator(mof_ctx* ctx, double dataScale, double weight, const double* x, double* y);  // vector_fields.cu
int vf_smooth_diagonal(mof_ctx* ctx, double* out);
int vf_triangle_field(mof_ctx* ctx, const double* coeffs, double* tfield);

// reorder.cu
int reorder_mesh(mof_ctx* ctx, int mode);
int reorder_gather(mof_ctx* ctx, int kind, const double* callerRows, int width, double* libraryRows);   // kind 0: per vertex, 1: per triangle
int reorder_scatter(mof_ctx* ctx, int kind, const double* libraryRows, int width, double* callerRows);

// setup_kernels.cu
int build_mesh_operators(mof_ctx* ctx);
int exclusive_scan_int(mof_ctx* ctx, const int* in, int* out, int n, int* total_out_device);
int reduce_sum(mof_ctx* ctx, const double* in, long long n, double* out_device);

// pcg_kernels.cu
// Jacobi-PCG, one right-hand side, matrix in the sliced layout (sliceBase has ceil(n/32)+1 entries).
int pcg_solve_sell(mof_ctx* ctx, int n, const int* sliceBase, const int* col, const double* val, const double* dinv, const double* b, double* x, bool zeroGuess,
                   double tol, int maxIters, int* itersOut, double* relresOut);
// Jacobi-PCG, six right-hand sides interleaved per row ([n][6]), matrix in CSR; x holds the initial guess.
int pcg_solve_csr6(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, const double* dinv, const double* b, double* x, bool zeroGuess,
                   double tol, int maxIters, int* itersOut, double* relresOut);
int extract_inverse_diagonal(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, double* dinv);
int time_spmv_sell(mof_ctx* ctx, int n, const int* sliceBase, const int* col, const double* val, const double* x, double* y, int reps, float* ms);
// CSR (device) -> sliced layout; allocates the three outputs. Used by the stand-alone solver entry.
int csr_to_sell(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, DBuf<int>& sliceBase, DBuf<int>& sCol, DBuf<double>& sVal);
// sliced layout -> CSR values/columns (rowptr is shared); for the debug taps.
int sell_to_csr(mof_ctx* ctx, int n, const int* rowptr, const int* sliceBase, const int* sCol, const double* sVal, int* col, double* val);

// y = A x on the sliced layout with per-CTA partials of x.y (the phase-1 code of the PCG kernel); *partials = CTA count.
int spmv_dot_launch(mof_ctx* ctx, int n, const int* sliceBase, const int* col, const double* val, const double* x, double* y, double* partial, int* partials);

// multigrid.cu
int mg_setup_mesh(mof_ctx* ctx);      // per mesh, both hierarchies; leaves one unusable (Jacobi-PCG stays) when the mesh does not fit
void mg_destroy(mof_ctx* ctx);
void mg_new_pair(mof_ctx* ctx);       // forget what the previous signal pair's systems left behind (warm starts of the set-up)
bool mg_flow_usable(const mof_ctx* ctx);
int mg_flow_update(mof_ctx* ctx);     // per flow system: coarse operators of the current wA
int mg_flow_solve(mof_ctx* ctx, double tol, int maxIters, int* itersOut, double* relresOut);   // wA fx = fb
bool mg_scalar_usable(const mof_ctx* ctx);
int mg_scalar_update(mof_ctx* ctx);   // per scalar system: coarse operators of the current sSys (sDinv = its inverse diagonal)
int mg_scalar_solve(mof_ctx* ctx, const double* b6, double* x6, double tol, int maxIters, int* itersOut, double* relresOut);  // x6 = initial guess
int mg_scalar_cycle(mof_ctx* ctx, const double* r6, double* z6);
bool mg_flow_try_update(mof_ctx* ctx);  // mg_flow_update for a matrix other than an alignment's; false: the hierarchy cannot take it
int mg_flow_cycle(mof_ctx* ctx, const double* r, double* z);  // one cycle of the flow hierarchy on the matrix of the last mg_flow_update
int mg_scalar_cheb(mof_ctx* ctx, const double* r6, double* z6, int degree, double lo);  // ... sharpened by `degree` Chebyshev steps around the cycle (fixed SPD operator)
int mg_scalar_smallest_eigenvalue(mof_ctx* ctx, int steps, double* lambdaMin);  // of (one cycle) x (current sSys), from the Lanczos tridiagonal of `steps` PCG iterations
int mg_time_kernel(mof_ctx* ctx, int which, int reps, float* ms, double* bytes);  // mof_time_kernel for the solver kernels  // z6 = one cycle applied to r6 (approximate inverse of the current sSys)

// dist.cu — one mesh over several GPUs: row blocks of the flow system, halo exchange, all-reduce (NCCL on ctx->stream)
int dist_unique_id(unsigned char* id128);
int dist_init(mof_ctx* ctx, int world, int rank, const unsigned char* id128);
void dist_destroy(mof_ctx* ctx);
// `kind`: 0 = the flow system (E edge rows, one value each), 1 = the smoothing systems (V vertex rows, six values each)
int dist_setup_mesh(mof_ctx* ctx);                       // per mesh: row blocks and halo index lists from the two patterns
bool dist_active(const mof_ctx* ctx);                    // a communicator exists and the mesh has been partitioned
int dist_world(const mof_ctx* ctx);
void dist_range(const mof_ctx* ctx, int kind, int* s0, int* s1, int* r0, int* r1);  // this rank's rows [r0,r1) (FLOW: = slices [s0,s1))
int dist_halo_f64(mof_ctx* ctx, int kind, double* vec);  // fills the entries of a full-length vector that my rows gather from other ranks
int dist_halo_f32(mof_ctx* ctx, int kind, float* vec);
int dist_allreduce_f64(mof_ctx* ctx, double* v, int count);
int dist_allreduce_f32(mof_ctx* ctx, float* v, int count);
int dist_allgather_rows(mof_ctx* ctx, int kind, double* vec);  // every rank's rows to every rank
// further partitions (any index set dealt in contiguous ranges; the indices outside the own range that this rank reads are flagged in ctx->itmp0[0..n))
int dist_add_partition(mof_ctx* ctx, int width, const int* rangeStart, int n, int* idOut);
void dist_clear_partitions(mof_ctx* ctx);
long long dist_partition_halo(const mof_ctx* ctx, int id);
int dist_halo_part_f32(mof_ctx* ctx, int id, float* vec);
int dist_allgather_part_f32(mof_ctx* ctx, int id, float* const* vecs, int count);
int dist_rank(const mof_ctx* ctx);
void dist_row_starts(const mof_ctx* ctx, int kind, int* out);  // world + 1 entries
int dist_p2p_setup(mof_ctx* ctx);   // after every partition exists: peer-memory windows for the halo exchanges (falls back to NCCL)
int dist_p2p_check(mof_ctx* ctx);   // after a solve: did every peer's halo arrive
int mg_dist_setup(mof_ctx* ctx);  // multigrid.cu: which coarse levels are dealt to the ranks, their cell ranges and halo lists (after dist_setup_mesh)

// vector_fields.cu — the Conformal and Connection bases (--vfMode 1|2): matrix-free block-Jacobi PCG
int vf_init(mof_ctx* ctx);                 // per signal pair, for ctx->params.vfMode / cMode (mode 0 releases the state)
void vf_destroy(mof_ctx* ctx);
bool vf_active(const mof_ctx* ctx);
bool vf_uses_scalar_hierarchy(const mof_ctx* ctx);  // Conformal with the two-cycle preconditioner: the flow solve re-values sSys and the scalar hierarchy
long long vf_unknowns(const mof_ctx* ctx);  // E (Whitney), 2V (Conformal) or 2T (Connection)
const double* vf_rhs(const mof_ctx* ctx);
const double* vf_solution(const mof_ctx* ctx);
int vf_update_flow(mof_ctx* ctx, double vfWeight);  // VectorField::UpdateOpticalFlow from ctx->dataD / dataRhs

// flow_kernels.cu
int dog_preprocess(mof_ctx* ctx);
int scalar_system_set(mof_ctx* ctx, double eps);  // sSys = M + eps S, its inverse diagonal and (if usable) the scalar hierarchy's coarse operators
void smooth_ahead_drain(mof_ctx* ctx);    // waits for a smoothing solve in flight and drops its result (before anything it reads changes)
void smooth_ahead_destroy(mof_ctx* ctx);
int update_flow(mof_ctx* ctx, double sWeight, double vfWeight);
int advect_vertices(mof_ctx* ctx, const double* in6, double lenA, double lenB, double* out6);
int advect_texels(mof_ctx* ctx, double alpha, int bilinear);
int advect_texels_frames(mof_ctx* ctx, int frames, int bilinear);  // ctx->texOut: [2][frames][W*H][3]
int time_walk_kernel(mof_ctx* ctx, int reps, float* ms);

// texprep_kernels.cu — the texture configuration's one-time preparation (MeshFlow.inl:158-467)
int subdivide_mesh(mof_ctx* ctx, double edgeLength, int* addedOut);                 // ctx->subXyz / subTri / subUv in place
int build_texture_map(mof_ctx* ctx, int W, int H, int padRadius, int* missesOut);  // ctx->triUV + edge transforms -> ctx->srcT / srcP
int sample_textures_to_vertices(mof_ctx* ctx, int bilinear, double* d_out6);        // ctx->tex[0|1] -> V x 6 (A rgb, B rgb)

}  // namespace mof
