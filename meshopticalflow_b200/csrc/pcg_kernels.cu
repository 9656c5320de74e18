// Jacobi-preconditioned conjugate gradients as ONE persistent cooperative kernel per solve.
//
// Replaces the reference's Eigen SimplicialLDLT / SimplicialLLT factor-and-solve
// (include/Misha/LinearSolvers.h:249-391; call sites VectorField.h:78-85, OpticalFlow.cpp:356-364,
// :828-840) on the solve path. The recurrence is the textbook PCG the reference itself carries,
// unused, in LinearSolvers.h:174-238 (SolvePreconditionedCG + DiagonalPreconditioner).
//
// Design (B200): the grid is sized to exactly fill the SMs (occupancy x SM count, co-resident), every
// CTA loops over its share of the rows, and the three phases of an iteration are separated by
// grid-wide barriers instead of kernel launches, so a solve of thousands of iterations is one launch
// with no host round trip. All scalars (alpha, beta, residual norms) are recomputed identically by
// every CTA from per-CTA partial sums written in a fixed order: bitwise deterministic, no atomics.
//
//   phase 1   q = A d, fused with d.q                          (the HBM-bound SpMV; see below)
//   phase 2   x += alpha d ; r -= alpha q ; fused r.Minv r and r.r
//   phase 3   d = Minv r + beta d
//
// SpMV, one right-hand side (flow system, ~11 nnz/row): the matrix is stored SLICED (SELL-32, see
// mof_internal.cuh): a warp owns 32 consecutive rows, lane = row, and entry j of all 32 rows is 32
// consecutive words, so every val/col load is one fully coalesced 256/128-byte request, touched exactly
// once (evict-first, the vectors keep the L2), with no shared-memory staging and no barrier; the
// loads of a row are independent, so each thread keeps a batch of them in flight before the dependent
// gathers of d[col]. Algorithmic bytes per launch: 12*nnz + 4*(n+1) + 16*n.
// SpMM, six right-hand sides (scalar smoothing, 7 nnz/row, CSR): vectors are [n][6] with the six
// channels of a vertex adjacent; one thread per (row, channel), so the six lanes of a row read val/col as
// a broadcast and the gathered 48 bytes as one coalesced piece, and every vector phase is a flat sweep.
//
// Tried and dropped (measured on B200, 3.1M rows): staging 256-row CSR tiles of val*d[col] in shared
// memory (117 us per SpMV) and feeding those tiles with cp.async.bulk/mbarrier rings of 2-4 stages
// (127-224 us: the ring's shared memory costs resident warps, and the kernel is bound by the latency of
// the dependent gathers, not by the streaming loads).
#include <cooperative_groups.h>

#include <cstdlib>

#include "mof_internal.cuh"

namespace cg = cooperative_groups;

namespace mof {

constexpr int PCG_T = 256;           // threads per CTA
constexpr int PCG_NW = PCG_T / 32;   // warps per CTA

template <int N>
struct PcgArgs {
    int n;
    const int* rowptr;     // CSR row pointers (N > 1) — unused for N == 1
    const int* sliceBase;  // sliced layout (N == 1)
    const int* col;
    const double* val;
    const double* dinv;
    const double* b;
    double* x;
    double* r;
    double* d;
    double* q;
    double* partial;  // 3 banks x gridDim x (3N)
    double* result;   // [0] iterations, [1] max relative residual (true), [2] converged
    double tol2;
    int maxIters;
    int zeroGuess;
};

// Sum of K per-thread values over the CTA, every thread gets the result; fixed order.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* sh) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double s = v[k];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sh[k * PCG_NW + w] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double s = 0;
#pragma unroll
        for (int i = 0; i < PCG_NW; i++) s += sh[k * PCG_NW + i];
        v[k] = s;
    }
    __syncthreads();
}

// Q per-thread values that belong to channel (global thread id % N) -> the Q*N per-channel CTA sums
// (v[q*N + j]), every thread gets them; fixed order. For N == 1 this is block_sum.
template <int N, int Q>
__device__ __forceinline__ void channel_sum(const double (&mine)[Q], double (&v)[Q * N], double* sh) {
    if (N == 1) {
        double t[Q * N];
#pragma unroll
        for (int q = 0; q < Q; q++) t[q] = mine[q];
        block_sum<Q * N>(t, sh);
#pragma unroll
        for (int q = 0; q < Q; q++) v[q] = t[q];
        return;
    }
#pragma unroll
    for (int q = 0; q < Q; q++) sh[q * PCG_T + threadIdx.x] = mine[q];
    __syncthreads();
    double s = 0;
    if (threadIdx.x < Q * N) {
        int q = threadIdx.x / N, j = threadIdx.x - q * N;
        // a CTA owns whole rows (thread t <-> local row t / N, channel t % N), so every channel is summed over the same
        // rows in the same order: identical right-hand sides give bitwise identical solutions
        for (int t = j; t < (PCG_T / N) * N; t += N) s += sh[q * PCG_T + t];
    }
    __syncthreads();
    if (threadIdx.x < Q * N) sh[threadIdx.x] = s;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < Q * N; k++) v[k] = sh[k];
    __syncthreads();
}

// CTA partial -> global; after the grid barrier every CTA folds all partials in the same order.
template <int K>
__device__ __forceinline__ void publish(const double (&v)[K], double* bank) {
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < K; k++) bank[(size_t)blockIdx.x * K + k] = v[k];
}
template <int K>
__device__ __forceinline__ void collect(const double* bank, double (&tot)[K], double* sh) {
#pragma unroll
    for (int k = 0; k < K; k++) tot[k] = 0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += PCG_T)
#pragma unroll
        for (int k = 0; k < K; k++) tot[k] += bank[(size_t)b * K + k];
    block_sum<K>(tot, sh);
}

template <int N>
__device__ __forceinline__ double pick(const double (&v)[N], int j) {
    double r = v[0];
#pragma unroll
    for (int k = 1; k < N; k++) r = j == k ? v[k] : r;
    return r;
}

// One right-hand side, sliced matrix: out = A in (mode 0, adds this thread's share of in.out to `dot`) or
// out = b - A in (mode 1). Warp = slice of 32 rows, lane = row; the row's entries are summed in column
// order, so the result does not depend on the launch geometry.
constexpr int SPMV_BATCH = 6;  // independent (val, col) pairs in flight per thread before the dependent gathers
__device__ __forceinline__ void spmv_sell(const int n, const int* __restrict__ sliceBase, const int* __restrict__ col, const double* __restrict__ val,
                                          const double* __restrict__ b, const double* __restrict__ in, double* __restrict__ out, int mode, double& dot) {
    const int lane = threadIdx.x & 31;
    const int slices = (n + 31) >> 5;
    const int warps = gridDim.x * PCG_NW;
    for (int s = blockIdx.x * PCG_NW + (threadIdx.x >> 5); s < slices; s += warps) {
        const int base = sliceBase[s];
        const int len = (sliceBase[s + 1] - base) >> 5;
        const double* v0 = val + (size_t)base + lane;
        const int* c0 = col + (size_t)base + lane;
        const int row = 32 * s + lane;
        double acc = 0;
        for (int j0 = 0; j0 < len; j0 += SPMV_BATCH) {
            double v[SPMV_BATCH], x[SPMV_BATCH];
            int c[SPMV_BATCH];
#pragma unroll
            for (int u = 0; u < SPMV_BATCH; u++) {
                bool ok = j0 + u < len;
                v[u] = ok ? __ldcs(v0 + 32 * (size_t)(j0 + u)) : 0.;
                c[u] = ok ? __ldcs(c0 + 32 * (size_t)(j0 + u)) : 0;
            }
#pragma unroll
            for (int u = 0; u < SPMV_BATCH; u++) x[u] = in[c[u]];
#pragma unroll
            for (int u = 0; u < SPMV_BATCH; u++)
                if (j0 + u < len) acc += v[u] * x[u];
        }
        if (row < n) {
            if (mode == 0) out[row] = acc, dot += in[row] * acc;
            else out[row] = b[row] - acc;
        }
    }
}

template <int N>
__global__ void __launch_bounds__(PCG_T, N == 1 ? 5 : 3) k_pcg(PcgArgs<N> a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[N == 1 ? 3 * PCG_NW : 3 * PCG_T];
    const int n = a.n;
    const size_t bankStride = (size_t)gridDim.x * 3 * N;
    double* bank0 = a.partial;
    double* bank1 = a.partial + bankStride;
    double* bank2 = a.partial + 2 * bankStride;
    // Flat element mapping: element i = row*N + channel. The sweep stride is a multiple of N, so a thread
    // always works on the same channel j.
    const size_t total = (size_t)n * N;
    constexpr int OWNED = (PCG_T / N) * N;  // threads of a CTA that own an element (whole rows per CTA)
    const size_t g = (size_t)blockIdx.x * OWNED + threadIdx.x;
    const size_t G = (size_t)gridDim.x * OWNED;
    const bool active = threadIdx.x < OWNED;
    const int j = (int)(threadIdx.x % N);

    double delta[N], bb[N], alpha[N], beta[N];
    bool frozen[N];

    auto spmv = [&](const double* __restrict__ in, double* __restrict__ out, int mode, double& dot) {
        if (N == 1) spmv_sell(n, a.sliceBase, a.col, a.val, a.b, in, out, mode, dot);
        else if (active)
            for (size_t i = g; i < total; i += G) {
                int row = (int)(i / N);
                double s = 0;
                for (int k = a.rowptr[row]; k < a.rowptr[row + 1]; k++) s += a.val[k] * in[(size_t)a.col[k] * N + j];
                if (mode == 0) out[i] = s, dot += in[i] * s;
                else out[i] = a.b[i] - s;
            }
    };

    // r = b - A x0 (or b), d = Minv r, delta = r.d, bb = b.b
    if (!a.zeroGuess) {
        double unused = 0;
        spmv(a.x, a.r, 1, unused);
        grid.sync();
    }
    {
        double mine[3] = {0, 0, 0}, acc[3 * N];
        if (active)
            for (size_t i = g; i < total; i += G) {
                double di = a.dinv[i / N], bv = a.b[i], rv;
                if (a.zeroGuess) rv = bv, a.r[i] = bv, a.x[i] = 0;
                else rv = a.r[i];
                double s = di * rv;
                a.d[i] = s;
                mine[0] += rv * s, mine[1] += rv * rv, mine[2] += bv * bv;
            }
        channel_sum<N, 3>(mine, acc, sh);
        publish<3 * N>(acc, bank0);
        grid.sync();
        collect<3 * N>(bank0, acc, sh);
#pragma unroll
        for (int c = 0; c < N; c++) delta[c] = acc[c], bb[c] = acc[2 * N + c], frozen[c] = !(acc[N + c] > a.tol2 * acc[2 * N + c]);
    }
    bool all = true;
#pragma unroll
    for (int c = 0; c < N; c++) all = all && frozen[c];

    int it = 0;
    while (!all && it < a.maxIters) {
        // phase 1: q = A d, d.q
        {
            double mine[1] = {0}, dq[N];
            spmv(a.d, a.q, 0, mine[0]);
            channel_sum<N, 1>(mine, dq, sh);
            publish<N>(dq, bank1);
            grid.sync();
            collect<N>(bank1, dq, sh);
#pragma unroll
            for (int c = 0; c < N; c++) alpha[c] = (!frozen[c] && dq[c] != 0) ? delta[c] / dq[c] : 0.;
        }
        // phase 2: x, r, fused r.Minv r and r.r
        double acc[2 * N];
        {
            double mine[2] = {0, 0};
            const double al = pick<N>(alpha, j);
            if (active)
                for (size_t i = g; i < total; i += G) {
                    double di = a.dinv[i / N];
                    double rv = a.r[i] - al * a.q[i];
                    a.x[i] += al * a.d[i];
                    a.r[i] = rv;
                    mine[0] += rv * (di * rv), mine[1] += rv * rv;
                }
            channel_sum<N, 2>(mine, acc, sh);
            publish<2 * N>(acc, bank2);
            grid.sync();
            collect<2 * N>(bank2, acc, sh);
        }
        it++;
        all = true;
#pragma unroll
        for (int c = 0; c < N; c++) {
            beta[c] = (!frozen[c] && delta[c] != 0) ? acc[c] / delta[c] : 0.;
            delta[c] = acc[c];
            if (!(acc[N + c] > a.tol2 * bb[c])) frozen[c] = true;
            all = all && frozen[c];
        }
        if (all) break;
        // phase 3: d = Minv r + beta d
        {
            const double be = pick<N>(beta, j);
            if (active)
                for (size_t i = g; i < total; i += G) a.d[i] = a.dinv[i / N] * a.r[i] + be * a.d[i];
        }
        grid.sync();
    }

    // true residual of the returned x: q = b - A x, max_j ||q_j|| / ||b_j||
    grid.sync();
    {
        double unused = 0;
        spmv(a.x, a.q, 1, unused);
    }
    grid.sync();
    {
        double mine[1] = {0}, acc[N];
        if (active)
            for (size_t i = g; i < total; i += G) {
                double v = a.q[i];
                mine[0] += v * v;
            }
        channel_sum<N, 1>(mine, acc, sh);
        publish<N>(acc, bank0);
        grid.sync();
        collect<N>(bank0, acc, sh);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            double worst = 0;
#pragma unroll
            for (int c = 0; c < N; c++) {
                double rel = bb[c] > 0 ? sqrt(acc[c] / bb[c]) : 0.;
                worst = rel > worst ? rel : worst;
            }
            a.result[0] = (double)it, a.result[1] = worst, a.result[2] = all ? 1. : 0.;
        }
    }
}

// Tuning knob (read once): cap on the CTAs per SM of the persistent grid.
static int ctas_per_sm_cap() {
    static int cap = [] {
        const char* e = getenv("MOF_PCG_CTAS_PER_SM");
        int v = e && *e ? atoi(e) : 0;
        return v > 0 ? v : 8;
    }();
    return cap;
}

template <int N>
static int pcg_grid(mof_ctx* ctx, int* grid) {
    int perSm = 0, sms = 0;
    MOF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_pcg<N>, PCG_T, 0));
    MOF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (perSm < 1) return fail(ctx, MOF_E_CUDA, "k_pcg does not fit on an SM");
    if (perSm > ctas_per_sm_cap()) perSm = ctas_per_sm_cap();
    *grid = perSm * sms;
    return MOF_OK;
}

template <int N>
static int launch_pcg(mof_ctx* ctx, PcgArgs<N>& args, int* iters, double* relres, bool* converged) {
    PcgWork& w = ctx->pcg;
    int grid = 0;
    MOF_TRY(pcg_grid<N>(ctx, &grid));
    MOF_CUDA(w.partial.reserve((size_t)grid * 3 * 6 * 3));
    MOF_CUDA(w.result.reserve(8));
    args.partial = w.partial.p, args.result = w.result.p;
    void* params[] = {&args};
#ifdef MOF_HOST_EMULATION  // CPU test tier, see tests/host_emulation: one CTA, or one OS thread per CTA
    (void)params;
    const PcgArgs<N> byValue = args;
    mof_emul::launch_cooperative(grid, PCG_T, [byValue] { k_pcg<N>(byValue); });
    cudaError_t e = cudaSuccess;
#else
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_pcg<N>, dim3(grid), dim3(PCG_T), params, 0, ctx->stream);
#endif
    ctx->stats.kernelLaunches++;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaLaunchCooperativeKernel(k_pcg)");
    double h[3];
    MOF_CUDA(cudaMemcpyAsync(h, w.result.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    *iters = (int)h[0], *relres = h[1], *converged = h[2] != 0;
    return MOF_OK;
}

template <int N>
static int pcg_run(mof_ctx* ctx, PcgArgs<N>& args, bool zeroGuess, double tol, int maxIters, int* itersOut, double* relresOut) {
    PcgWork& w = ctx->pcg;
    size_t len = (size_t)args.n * N;
    if (w.r.n < len) {
        MOF_CUDA(w.r.alloc(len));
        MOF_CUDA(w.d.alloc(len));
        MOF_CUDA(w.q.alloc(len));
    }
    args.r = w.r.p, args.d = w.d.p, args.q = w.q.p;
    args.tol2 = tol * tol, args.zeroGuess = zeroGuess ? 1 : 0;
    int total = 0, iters = 0;
    double relres = 0;
    bool converged = false;
    // The recurrence residual can drift from the true one over thousands of iterations: when the
    // true residual of the returned x misses the tolerance, restart from x (at most a few times).
    for (int attempt = 0; attempt < 6; attempt++) {
        args.maxIters = maxIters - total;
        MOF_TRY(launch_pcg<N>(ctx, args, &iters, &relres, &converged));
        total += iters;
        if (relres <= tol * 1.0001 || total >= maxIters) break;
        if (!converged) break;
        args.zeroGuess = 0;
    }
    *itersOut = total, *relresOut = relres;
    if (!(relres <= tol * 1.0001)) ctx->stats.solvesAboveTolerance++;  // accepted below MOF_ACCEPT_RELRES, or about to fail: visible either way
    if (!(relres <= tol * 1.0001) && (total >= maxIters || !(relres <= MOF_ACCEPT_RELRES))) {
        char msg[160];
        snprintf(msg, sizeof(msg), "[ERROR] PCG did not reach %g in %d iterations (relative residual %g)", tol, total, relres);
        return fail(ctx, MOF_E_NOCONVERGE, msg);
    }
    return MOF_OK;
}

int pcg_solve_sell(mof_ctx* ctx, int n, const int* sliceBase, const int* col, const double* val, const double* dinv, const double* b, double* x, bool zeroGuess,
                   double tol, int maxIters, int* itersOut, double* relresOut) {
    PcgArgs<1> args = {};
    args.n = n, args.sliceBase = sliceBase, args.col = col, args.val = val, args.dinv = dinv, args.b = b, args.x = x;
    return pcg_run<1>(ctx, args, zeroGuess, tol, maxIters, itersOut, relresOut);
}

int pcg_solve_csr6(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, const double* dinv, const double* b, double* x, bool zeroGuess,
                   double tol, int maxIters, int* itersOut, double* relresOut) {
    PcgArgs<6> args = {};
    args.n = n, args.rowptr = rowptr, args.col = col, args.val = val, args.dinv = dinv, args.b = b, args.x = x;
    return pcg_run<6>(ctx, args, zeroGuess, tol, maxIters, itersOut, relresOut);
}

// DiagonalPreconditioner::set, LinearSolvers.h:88-104.
__global__ void k_inverse_diagonal(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int n, double* __restrict__ dinv) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double s = 0;
    for (int k = rowptr[row]; k < rowptr[row + 1]; k++)
        if (col[k] == row) s += val[k];
    dinv[row] = 1. / s;
}

int extract_inverse_diagonal(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, double* dinv) {
    MOF_LAUNCH(k_inverse_diagonal, blocks_for(n, 256), 256, 0, rowptr, col, val, n, dinv);
    return MOF_OK;
}

// The phase-1 code on its own, for the roofline line of bench.py and for ncu: y = A x fused with x.y,
// same code, same grid as inside k_pcg<1>.
__global__ void __launch_bounds__(PCG_T, 5) k_spmv_dot(int n, const int* __restrict__ sliceBase, const int* __restrict__ col, const double* __restrict__ val,
                                                       const double* __restrict__ x, double* __restrict__ y, double* __restrict__ partial) {
    pdl_wait();
    __shared__ double sh[PCG_NW];
    double dot[1] = {0};
    spmv_sell(n, sliceBase, col, val, nullptr, x, y, 0, dot[0]);
    block_sum<1>(dot, sh);
    publish<1>(dot, partial);
}

int spmv_dot_launch(mof_ctx* ctx, int n, const int* sliceBase, const int* col, const double* val, const double* x, double* y, double* partial, int* partials) {
    int& grid = ctx->pcg.gridBlocks;  // per context (per device), computed once
    if (!grid) MOF_TRY(pcg_grid<1>(ctx, &grid));
    MOF_LAUNCH_PDL(k_spmv_dot, grid, PCG_T, 0, n, sliceBase, col, val, x, y, partial);
    *partials = grid;
    return MOF_OK;
}

int time_spmv_sell(mof_ctx* ctx, int n, const int* sliceBase, const int* col, const double* val, const double* x, double* y, int reps, float* ms) {
    int grid = 0;
    MOF_TRY(pcg_grid<1>(ctx, &grid));
    MOF_CUDA(ctx->pcg.partial.reserve((size_t)grid * 3 * 6 * 3));
    for (int i = 0; i < 3; i++) MOF_LAUNCH(k_spmv_dot, grid, PCG_T, 0, n, sliceBase, col, val, x, y, ctx->pcg.partial.p);
    MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < reps; i++) MOF_LAUNCH(k_spmv_dot, grid, PCG_T, 0, n, sliceBase, col, val, x, y, ctx->pcg.partial.p);
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    float t = 0;
    MOF_CUDA(cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
    *ms = t / reps;
    return MOF_OK;
}

}  // namespace mof
