// Jacobi-preconditioned conjugate gradients as ONE persistent cooperative kernel per solve.
//
// Replaces the reference's Eigen SimplicialLDLT / SimplicialLLT factor-and-solve
// (include/Misha/LinearSolvers.h:249-391; call sites VectorField.h:78-85, OpticalFlow.cpp:356-364,
// :828-840) on the solve path. The recurrence is the textbook PCG the reference itself carries,
// unused, in LinearSolvers.h:174-238 (SolvePreconditionedCG + DiagonalPreconditioner).
//
// Design (B200): the grid is sized to exactly fill the SMs (occupancy x SM count, co-resident), every
// CTA loops over row tiles, and the three phases of an iteration are separated by grid-wide barriers
// instead of kernel launches, so a solve of thousands of iterations is one launch with no host
// round trip. All scalars (alpha, beta, residual norms) are recomputed identically by every CTA from
// per-CTA partial sums written in a fixed order: bitwise deterministic, no atomics.
//
//   phase 1   q = A d, fused with d.q                          (the HBM-bound SpMV; see below)
//   phase 2   x += alpha d ; r -= alpha q ; fused r.Minv r and r.r
//   phase 3   d = Minv r + beta d
//
// SpMV, one right-hand side (flow system, ~11 nnz/row): a CTA stages the products val[k]*d[col[k]]
// of a 256-row tile in shared memory with fully coalesced streaming loads of val/col (each array is
// touched exactly once), then one thread per row sums its segment in row order. Algorithmic bytes per
// launch: 12*nnz + 4*(n+1) + 16*n.
// SpMV, six right-hand sides (scalar smoothing, 7 nnz/row): one thread per row, the six channels of
// a vertex are adjacent in memory so the matrix is read once for all six.
#include <cooperative_groups.h>

#include "mof_internal.cuh"

namespace cg = cooperative_groups;

namespace mof {

constexpr int PCG_T = 256;           // threads per CTA
constexpr int PCG_NW = PCG_T / 32;   // warps per CTA
constexpr int TILE_ROWS = PCG_T;     // rows per SpMV tile
constexpr int PROD_CAP = 4096;       // staged products per tile (32 KB)

template <int N>
struct PcgArgs {
    int n;
    const int* rowptr;
    const int* col;
    const double* val;
    const double* dinv;
    const double* b;
    double* x;
    double* r;
    double* d;
    double* q;
    double* partial;  // 3 banks x gridDim x (2N)
    double* result;   // [0] iterations, [1] max relative residual (true), [2] converged
    double tol2;
    int maxIters;
    int zeroGuess;
};

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

// CTA partial -> global; after the grid barrier every CTA folds all partials in the same order.
template <int K>
__device__ __forceinline__ void publish(double (&v)[K], double* bank, double* sh) {
    block_sum<K>(v, sh);
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

// out = A in (mode 0, dot += in.out) or out = b - A in (mode 1).
template <int N>
__device__ __forceinline__ void spmv(const PcgArgs<N>& a, const double* __restrict__ in, double* __restrict__ out, int mode, double (&dot)[N], double* prod) {
    const int n = a.n;
    if (N == 1) {
        int tiles = (n + TILE_ROWS - 1) / TILE_ROWS;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            int r0 = tile * TILE_ROWS, r1 = min(n, r0 + TILE_ROWS);
            int k0 = a.rowptr[r0], k1 = a.rowptr[r1];
            int row = r0 + threadIdx.x;
            double s = 0;
            if (k1 - k0 <= PROD_CAP) {
                for (int k = k0 + threadIdx.x; k < k1; k += PCG_T) prod[k - k0] = a.val[k] * in[a.col[k]];
                __syncthreads();
                if (row < r1) {
                    int kb = a.rowptr[row] - k0, ke = a.rowptr[row + 1] - k0;
                    for (int k = kb; k < ke; k++) s += prod[k];
                }
                __syncthreads();
            } else if (row < r1) {
                for (int k = a.rowptr[row]; k < a.rowptr[row + 1]; k++) s += a.val[k] * in[a.col[k]];
            }
            if (row < r1) {
                if (mode == 0) out[row] = s, dot[0] += in[row] * s;
                else out[row] = a.b[row] - s;
            }
        }
    } else {
        for (int row = blockIdx.x * PCG_T + threadIdx.x; row < n; row += gridDim.x * PCG_T) {
            double s[N];
#pragma unroll
            for (int j = 0; j < N; j++) s[j] = 0;
            for (int k = a.rowptr[row]; k < a.rowptr[row + 1]; k++) {
                double v = a.val[k];
                const double* src = in + (size_t)a.col[k] * N;
#pragma unroll
                for (int j = 0; j < N; j++) s[j] += v * src[j];
            }
#pragma unroll
            for (int j = 0; j < N; j++) {
                if (mode == 0) out[(size_t)row * N + j] = s[j], dot[j] += in[(size_t)row * N + j] * s[j];
                else out[(size_t)row * N + j] = a.b[(size_t)row * N + j] - s[j];
            }
        }
    }
}

template <int N>
__global__ void __launch_bounds__(PCG_T) k_pcg(PcgArgs<N> a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double prod[N == 1 ? PROD_CAP : 1];
    __shared__ double sh[3 * N * PCG_NW];
    const int n = a.n;
    const size_t bankStride = (size_t)gridDim.x * 3 * N;
    double* bank0 = a.partial;
    double* bank1 = a.partial + bankStride;
    double* bank2 = a.partial + 2 * bankStride;
    const int gtid = blockIdx.x * PCG_T + threadIdx.x, gsz = gridDim.x * PCG_T;

    double delta[N], bb[N], alpha[N], beta[N];
    bool frozen[N];
    double dummy[N];

    // r = b - A x0 (or b), d = Minv r, delta = r.d, bb = b.b
    if (!a.zeroGuess) {
        spmv<N>(a, a.x, a.r, 1, dummy, prod);
    }
    {
        double acc[3 * N];
#pragma unroll
        for (int k = 0; k < 3 * N; k++) acc[k] = 0;
        for (int row = gtid; row < n; row += gsz) {
            double di = a.dinv[row];
#pragma unroll
            for (int j = 0; j < N; j++) {
                size_t i = (size_t)row * N + j;
                double bv = a.b[i], rv;
                if (a.zeroGuess) rv = bv, a.r[i] = bv, a.x[i] = 0;
                else rv = a.r[i];
                double s = di * rv;
                a.d[i] = s;
                acc[j] += rv * s, acc[N + j] += rv * rv, acc[2 * N + j] += bv * bv;
            }
        }
        publish<3 * N>(acc, bank0, sh);
        grid.sync();
        collect<3 * N>(bank0, acc, sh);
#pragma unroll
        for (int j = 0; j < N; j++) delta[j] = acc[j], bb[j] = acc[2 * N + j], frozen[j] = !(acc[N + j] > a.tol2 * acc[2 * N + j]);
    }
    bool all = true;
#pragma unroll
    for (int j = 0; j < N; j++) all = all && frozen[j];

    int it = 0;
    while (!all && it < a.maxIters) {
        // phase 1: q = A d, d.q
        double dq[N];
#pragma unroll
        for (int j = 0; j < N; j++) dq[j] = 0;
        spmv<N>(a, a.d, a.q, 0, dq, prod);
        publish<N>(dq, bank1, sh);
        grid.sync();
        collect<N>(bank1, dq, sh);
#pragma unroll
        for (int j = 0; j < N; j++) alpha[j] = (!frozen[j] && dq[j] != 0) ? delta[j] / dq[j] : 0.;

        // phase 2: x, r, fused r.Minv r and r.r
        double acc[2 * N];
#pragma unroll
        for (int k = 0; k < 2 * N; k++) acc[k] = 0;
        for (int row = gtid; row < n; row += gsz) {
            double di = a.dinv[row];
#pragma unroll
            for (int j = 0; j < N; j++) {
                size_t i = (size_t)row * N + j;
                double rv = a.r[i] - alpha[j] * a.q[i];
                a.x[i] += alpha[j] * a.d[i];
                a.r[i] = rv;
                acc[j] += rv * (di * rv), acc[N + j] += rv * rv;
            }
        }
        publish<2 * N>(acc, bank2, sh);
        grid.sync();
        collect<2 * N>(bank2, acc, sh);
        it++;
        all = true;
#pragma unroll
        for (int j = 0; j < N; j++) {
            beta[j] = (!frozen[j] && delta[j] != 0) ? acc[j] / delta[j] : 0.;
            delta[j] = acc[j];
            if (!(acc[N + j] > a.tol2 * bb[j])) frozen[j] = true;
            all = all && frozen[j];
        }
        if (all) break;

        // phase 3: d = Minv r + beta d
        for (int row = gtid; row < n; row += gsz) {
            double di = a.dinv[row];
#pragma unroll
            for (int j = 0; j < N; j++) {
                size_t i = (size_t)row * N + j;
                a.d[i] = di * a.r[i] + beta[j] * a.d[i];
            }
        }
        grid.sync();
    }

    // true residual of the returned x: q = b - A x, max_j ||q_j|| / ||b_j||
    grid.sync();
    spmv<N>(a, a.x, a.q, 1, dummy, prod);
    grid.sync();
    {
        double acc[N];
#pragma unroll
        for (int j = 0; j < N; j++) acc[j] = 0;
        for (int row = gtid; row < n; row += gsz)
#pragma unroll
            for (int j = 0; j < N; j++) {
                double v = a.q[(size_t)row * N + j];
                acc[j] += v * v;
            }
        publish<N>(acc, bank0, sh);
        grid.sync();
        collect<N>(bank0, acc, sh);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            double worst = 0;
#pragma unroll
            for (int j = 0; j < N; j++) {
                double rel = bb[j] > 0 ? sqrt(acc[j] / bb[j]) : 0.;
                worst = rel > worst ? rel : worst;
            }
            a.result[0] = (double)it, a.result[1] = worst, a.result[2] = all ? 1. : 0.;
        }
    }
}

template <int N>
static int launch_pcg(mof_ctx* ctx, PcgArgs<N>& args, int* iters, double* relres, bool* converged) {
    PcgWork& w = ctx->pcg;
    int perSm = 0, sms = 0;
    MOF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_pcg<N>, PCG_T, 0));
    MOF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (perSm < 1) return fail(ctx, MOF_E_CUDA, "k_pcg does not fit on an SM");
    if (perSm > 4) perSm = 4;
    int grid = perSm * sms;
    MOF_CUDA(w.partial.reserve((size_t)grid * 3 * 6 * 3));
    MOF_CUDA(w.result.reserve(8));
    args.partial = w.partial.p, args.result = w.result.p;
    void* params[] = {&args};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_pcg<N>, dim3(grid), dim3(PCG_T), params, 0, ctx->stream);
    ctx->stats.kernelLaunches++;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaLaunchCooperativeKernel(k_pcg)");
    double h[3];
    MOF_CUDA(cudaMemcpyAsync(h, w.result.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    *iters = (int)h[0], *relres = h[1], *converged = h[2] != 0;
    return MOF_OK;
}

template <int N>
static int pcg_solve_n(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, const double* dinv, const double* b, double* x,
                       bool zeroGuess, double tol, int maxIters, int* itersOut, double* relresOut) {
    PcgWork& w = ctx->pcg;
    size_t len = (size_t)n * N;
    if (w.r.n < len) {
        MOF_CUDA(w.r.alloc(len));
        MOF_CUDA(w.d.alloc(len));
        MOF_CUDA(w.q.alloc(len));
    }
    PcgArgs<N> args;
    args.n = n, args.rowptr = rowptr, args.col = col, args.val = val, args.dinv = dinv, args.b = b, args.x = x;
    args.r = w.r.p, args.d = w.d.p, args.q = w.q.p;
    args.tol2 = tol * tol, args.maxIters = maxIters, args.zeroGuess = zeroGuess ? 1 : 0;
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
    if (!(relres <= tol * 1.0001) && (total >= maxIters || !(relres <= 1e-4))) {
        char msg[160];
        snprintf(msg, sizeof(msg), "[ERROR] PCG did not reach %g in %d iterations (relative residual %g)", tol, total, relres);
        return fail(ctx, MOF_E_NOCONVERGE, msg);
    }
    return MOF_OK;
}

int pcg_solve(mof_ctx* ctx, int n, long long nnz, const int* rowptr, const int* col, const double* val, const double* dinv, const double* b, double* x,
              int nrhs, bool zeroGuess, double tol, int maxIters, int* itersOut, double* relresOut) {
    (void)nnz;
    if (nrhs == 1) return pcg_solve_n<1>(ctx, n, rowptr, col, val, dinv, b, x, zeroGuess, tol, maxIters, itersOut, relresOut);
    if (nrhs == 6) return pcg_solve_n<6>(ctx, n, rowptr, col, val, dinv, b, x, zeroGuess, tol, maxIters, itersOut, relresOut);
    return fail(ctx, MOF_E_INVALID, "pcg_solve: nrhs must be 1 or 6");
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

// The phase-1 kernel on its own, for the roofline line of bench.py and for ncu: y = A x fused with
// x.y, same tiles, same grid as inside k_pcg.
__global__ void __launch_bounds__(PCG_T) k_spmv_dot(PcgArgs<1> a, const double* __restrict__ x, double* __restrict__ y) {
    __shared__ double prod[PROD_CAP];
    __shared__ double sh[PCG_NW];
    double dot[1] = {0};
    spmv<1>(a, x, y, 0, dot, prod);
    publish<1>(dot, a.partial, sh);
}

int time_spmv(mof_ctx* ctx, int n, long long nnz, const int* rowptr, const int* col, const double* val, const double* x, double* y, int reps, float* ms) {
    (void)nnz;
    int perSm = 0, sms = 0;
    MOF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_pcg<1>, PCG_T, 0));
    MOF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (perSm > 4) perSm = 4;
    int grid = perSm * sms;
    MOF_CUDA(ctx->pcg.partial.reserve((size_t)grid * 3 * 6 * 3));
    PcgArgs<1> args = {};
    args.n = n, args.rowptr = rowptr, args.col = col, args.val = val, args.partial = ctx->pcg.partial.p;
    for (int i = 0; i < 3; i++) MOF_LAUNCH(k_spmv_dot, grid, PCG_T, 0, args, x, y);
    MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < reps; i++) MOF_LAUNCH(k_spmv_dot, grid, PCG_T, 0, args, x, y);
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    float t = 0;
    MOF_CUDA(cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1));
    *ms = t / reps;
    return MOF_OK;
}

}  // namespace mof
