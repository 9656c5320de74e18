// The lowest eigenpairs of the vector Laplacian of a basis, S x = lambda M x (the reference's `Spectrum` tool).
//
// Reference: ComputeSpectrum (include/Src/VectorLaplacianSpectrum.inl:5-39) forms M = R (g area) P by two sparse products, factorises
// S - 1e-8 M (Eigen::SimplicialLDLT) and runs ARPACK's shift-invert Lanczos on it (EigenvalueSolver.h:177-219), then prolongs the
// eigenvectors to per-triangle fields (P x). Here there is no factorisation and no Krylov recurrence through exact solves: a block
// method that only APPLIES the operators — LOBPCG (locally optimal block preconditioned conjugate gradients, Knyazev 2001) — on
// the operators the alignment path already has on the device:
//   Whitney     S = ctx->wS (sliced layout), M assembled once on the same pattern by the flow assembly with D_t = g_t area_t;
//   Conformal / Connection   S and M applied matrix-free by the kernels of vector_fields.cu ((0, 1) and (1, 0) of s P^T D P + w S).
// Per iteration, for a block of m = count + guard vectors (guard = max(4, count/2 + 2), m <= 32): the residuals R = S X - M X diag(theta), W = T R (T = the inverse
// diagonal of S), a Rayleigh-Ritz step on span[X, W, P] — two Gram matrices of 3m x 3m by a tiled kernel with per-CTA partials
// folded in a fixed order (deterministic), the dense generalised eigenproblem of that size on the host (Cholesky + cyclic Jacobi,
// <= 96 x 96), and the new X, P, S X, M X, S P, M P as block combinations (one kernel). Vectors are column-major (each column
// contiguous) so every single-vector kernel of the library applies to a column as it is.
// Converged when every wanted pair has ||S x - theta M x|| <= tol * (||S x|| + max(|theta|, theta_max / 1000) ||M x||).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mof_internal.cuh"

namespace mof {

namespace {

constexpr int B = 256;
constexpr int MAXM = 32;          // block size limit: count + guard <= 32
constexpr int GRAM_ROWS = 64;     // rows per tile of the Gram kernel
constexpr int GRAM_CTAS = 296;    // per-CTA partials of the Gram kernel (2 per SM)

struct Ops {
    int mode = 0;
    long long n = 0;
    bool cycle = false;  // Whitney on a mesh with a flow hierarchy: T = one multigrid cycle on S + tau M instead of the inverse diagonal
    double tau = 0;
    bool twoCycle = false;  // Conformal on a mesh with a scalar hierarchy: T = 2 eps^2 C M C (below) instead of the inverse diagonal
    int chebDegree = 1;
    double chebLo = 0.1;
    ScopedBuf<double> wM, tinv, r6, z6;
};

// y = A x, A in the sliced layout (padding entries carry value 0 and the row's own column).
__global__ void k_sell_apply(int n, const int* __restrict__ sliceBase, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x,
                             double* __restrict__ y) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int s = r >> 5;
    const int len = (sliceBase[s + 1] - sliceBase[s]) >> 5;
    double acc = 0;
    for (int j = 0; j < len; j++) {
        const size_t k = sell_pos(sliceBase, r, j);
        acc += val[k] * x[col[k]];
    }
    y[r] = acc;
}
__global__ void k_sell_inverse_diagonal(int n, const int* __restrict__ sliceBase, const int* __restrict__ col, const double* __restrict__ val, double* __restrict__ out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int s = r >> 5;
    const int len = (sliceBase[s + 1] - sliceBase[s]) >> 5;
    double d = 0;
    for (int j = 0; j < len; j++) {
        const size_t k = sell_pos(sliceBase, r, j);
        if (col[k] == r) d += val[k];
    }
    out[r] = d > 0 ? 1. / d : 0.;
}
// A = S + tau M on the padded sliced layout, with its inverse diagonal (the matrix the multigrid preconditioner is built for).
__global__ void k_shifted_operator(int n, const int* __restrict__ sliceBase, const int* __restrict__ col, const double* __restrict__ S, const double* __restrict__ M,
                                   double tau, double* __restrict__ A, double* __restrict__ dinv) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int s = r >> 5;
    const int len = (sliceBase[s + 1] - sliceBase[s]) >> 5;
    double d = 0;
    for (int j = 0; j < len; j++) {
        const size_t k = sell_pos(sliceBase, r, j);
        const double v = S[k] + tau * M[k];
        A[k] = v;
        if (col[k] == r) d += v;
    }
    dinv[r] = d > 0 ? 1. / d : 0.;
}
__global__ void k_invert_positive(long long n, double* __restrict__ v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i] > 0 ? 1. / v[i] : 0.;
}
__global__ void k_random_block(long long n, int m, double* __restrict__ x) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * m) return;
    unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    h ^= h >> 32, h *= 0xD6E8FEB86659FD93ull, h ^= h >> 32, h *= 0xD6E8FEB86659FD93ull, h ^= h >> 32;
    x[i] = (double)(h >> 11) * (1. / 9007199254740992.) - 0.5;
}
// R[:, j] = SX[:, j] - theta[j] MX[:, j];  W[:, j] = tinv .* R[:, j]
__global__ void k_residual_block(long long n, int m, const double* __restrict__ SX, const double* __restrict__ MX, const double* __restrict__ theta,
                                 const double* __restrict__ tinv, double* __restrict__ R, double* __restrict__ W) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * m) return;
    const int j = (int)(i / n);
    const long long r = i - (long long)j * n;
    const double v = SX[i] - theta[j] * MX[i];
    R[i] = v, W[i] = tinv[r] * v;
}

// Conformal basis: constants of either potential are in the null space of S AND of M (P maps them to the zero field); the iteration runs in
// their complement. One CTA per (column, half): subtracts the half's mean.
__global__ void __launch_bounds__(B) k_remove_half_means(long long half, double* __restrict__ X) {
    __shared__ double sh[B];
    double* x = X + (size_t)blockIdx.x * half;
    double s = 0;
    for (long long i = threadIdx.x; i < half; i += B) s += x[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    const double mean = sh[0] / (double)half;
    for (long long i = threadIdx.x; i < half; i += B) x[i] -= mean;
}

// Conformal basis, preconditioner T = 2 eps^2 C M C on each half (C ~ (M + eps K)^-1 by the scalar hierarchy, eps = 1/tau: the inverse of
// (K + tau M) M^-1 (K + tau M) / 2 = S + tau K + tau^2 M / 2, cf. the flow solve's two-cycle preconditioner in vector_fields.cu). The scalar
// hierarchy carries six channels: THREE columns of the block (two halves each) ride in one application.
__global__ void k_conformal_pack3(const double* __restrict__ R, long long n, int V, int cols, double* __restrict__ r6) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double* d = r6 + 6 * (size_t)v;
    for (int k = 0; k < 3; k++) {
        d[2 * k] = k < cols ? R[(size_t)k * n + v] : 0.;
        d[2 * k + 1] = k < cols ? R[(size_t)k * n + V + v] : 0.;
    }
}
__global__ void k_conformal_weight6(const double* __restrict__ z6, const double* __restrict__ m0, int V, double* __restrict__ r6) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6ll * V) return;
    r6[i] = m0[i / 6] * z6[i];
}
__global__ void k_conformal_unpack3(const double* __restrict__ z6, double kappa, long long n, int V, int cols, double* __restrict__ W) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const double* d = z6 + 6 * (size_t)v;
    for (int k = 0; k < cols; k++) W[(size_t)k * n + v] = kappa * d[2 * k], W[(size_t)k * n + V + v] = kappa * d[2 * k + 1];
}

// G[i][j] = sum_r A[r + i n] B[r + j n] for i < ka, j < kb: per-CTA partials (tiles of GRAM_ROWS rows through shared memory), folded below.
__global__ void __launch_bounds__(B) k_gram_partial(const double* __restrict__ A, const double* __restrict__ Bm, long long n, int ka, int kb, double* __restrict__ partial) {
    __shared__ double sa[MAXM][GRAM_ROWS + 1], sb[MAXM][GRAM_ROWS + 1];
    constexpr int PER = (MAXM * MAXM + B - 1) / B;
    double acc[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) acc[q] = 0;
    const int outs = ka * kb;
    for (long long r0 = (long long)blockIdx.x * GRAM_ROWS; r0 < n; r0 += (long long)gridDim.x * GRAM_ROWS) {
        const int rows = (int)min((long long)GRAM_ROWS, n - r0);
        for (int q = threadIdx.x; q < ka * GRAM_ROWS; q += B) {
            const int c = q / GRAM_ROWS, r = q - c * GRAM_ROWS;
            sa[c][r] = r < rows ? A[r0 + r + (long long)c * n] : 0.;
        }
        for (int q = threadIdx.x; q < kb * GRAM_ROWS; q += B) {
            const int c = q / GRAM_ROWS, r = q - c * GRAM_ROWS;
            sb[c][r] = r < rows ? Bm[r0 + r + (long long)c * n] : 0.;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int o = threadIdx.x + q * B;
            if (o < outs) {
                const int i = o / kb, j = o - i * kb;
                double s = 0;
                for (int r = 0; r < GRAM_ROWS; r++) s += sa[i][r] * sb[j][r];
                acc[q] += s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const int o = threadIdx.x + q * B;
        if (o < outs) partial[(size_t)blockIdx.x * outs + o] = acc[q];
    }
}
__global__ void k_gram_fold(const double* __restrict__ partial, int np, int outs, double* __restrict__ out) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= outs) return;
    double s = 0;
    for (int p = 0; p < np; p++) s += partial[(size_t)p * outs + o];
    out[o] = s;
}
// out[:, j] = sum_b sum_i in_b[:, i] C[(b kIn + i) * kOut + j]: up to three input blocks of kIn columns each; out may not alias an input.
struct Blocks3 {
    const double* in[3];
    int count;
};
__global__ void __launch_bounds__(B) k_combine(Blocks3 blk, long long n, int kIn, int kOut, const double* __restrict__ C, double* __restrict__ out) {
    __shared__ double sc[3 * MAXM * MAXM];
    const int rowsC = blk.count * kIn;
    for (int q = threadIdx.x; q < rowsC * kOut; q += B) sc[q] = C[q];
    __syncthreads();
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double acc[MAXM];
#pragma unroll
    for (int j = 0; j < MAXM; j++) acc[j] = 0;
    for (int b = 0; b < blk.count; b++)
        for (int i = 0; i < kIn; i++) {
            const double v = blk.in[b][r + (long long)i * n];
            const double* c = sc + (size_t)(b * kIn + i) * kOut;
#pragma unroll
            for (int j = 0; j < MAXM; j++)
                if (j < kOut) acc[j] += v * c[j];
        }
#pragma unroll
    for (int j = 0; j < MAXM; j++)
        if (j < kOut) out[r + (long long)j * n] = acc[j];
}

// ------------------------------------------------------------------------------------------------ host: small dense algebra

// Cholesky G = L L^T in place (lower); false if G is not positive definite to working precision.
bool cholesky(std::vector<double>& g, int k) {
    for (int j = 0; j < k; j++) {
        double d = g[j * k + j];
        for (int p = 0; p < j; p++) d -= g[j * k + p] * g[j * k + p];
        if (!(d > 1e-14 * std::fabs(g[j * k + j])) || !std::isfinite(d)) return false;
        d = std::sqrt(d);
        g[j * k + j] = d;
        for (int i = j + 1; i < k; i++) {
            double s = g[i * k + j];
            for (int p = 0; p < j; p++) s -= g[i * k + p] * g[j * k + p];
            g[i * k + j] = s / d;
        }
    }
    for (int i = 0; i < k; i++)
        for (int j = i + 1; j < k; j++) g[i * k + j] = 0;
    return true;
}
// Eigen-decomposition of a symmetric k x k matrix by cyclic Jacobi rotations: a -> eigenvalues on its diagonal, v = eigenvectors (columns).
void jacobi_eigen(std::vector<double>& a, int k, std::vector<double>& v) {
    v.assign((size_t)k * k, 0.);
    for (int i = 0; i < k; i++) v[i * k + i] = 1;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0, diag = 0;
        for (int i = 0; i < k; i++)
            for (int j = 0; j < k; j++) (i == j ? diag : off) += a[i * k + j] * a[i * k + j];
        if (off <= 1e-30 * diag) break;
        for (int p = 0; p < k; p++)
            for (int q = p + 1; q < k; q++) {
                const double apq = a[p * k + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double tau = (a[q * k + q] - a[p * k + p]) / (2 * apq);
                const double t = (tau >= 0 ? 1. : -1.) / (std::fabs(tau) + std::sqrt(1 + tau * tau));
                const double c = 1 / std::sqrt(1 + t * t), s = t * c;
                for (int i = 0; i < k; i++) {
                    const double aip = a[i * k + p], aiq = a[i * k + q];
                    a[i * k + p] = c * aip - s * aiq, a[i * k + q] = s * aip + c * aiq;
                }
                for (int i = 0; i < k; i++) {
                    const double api = a[p * k + i], aqi = a[q * k + i];
                    a[p * k + i] = c * api - s * aqi, a[q * k + i] = s * api + c * aqi;
                }
                for (int i = 0; i < k; i++) {
                    const double vip = v[i * k + p], viq = v[i * k + q];
                    v[i * k + p] = c * vip - s * viq, v[i * k + q] = s * vip + c * viq;
                }
            }
    }
}
// The `want` lowest eigenpairs of GA c = theta GM c (k x k, symmetric, GM positive definite): C [k][want] (GM-orthonormal columns), theta ascending.
bool small_eigen(const std::vector<double>& GA, const std::vector<double>& GM, int k, int want, std::vector<double>& C, std::vector<double>& theta) {
    std::vector<double> L = GM;
    // scale to unit diagonal first: the blocks of the basis have very different norms
    std::vector<double> d(k);
    for (int i = 0; i < k; i++) {
        if (!(GM[i * k + i] > 0)) return false;
        d[i] = 1 / std::sqrt(GM[i * k + i]);
    }
    std::vector<double> A(k * k);
    for (int i = 0; i < k; i++)
        for (int j = 0; j < k; j++) L[i * k + j] = GM[i * k + j] * d[i] * d[j], A[i * k + j] = GA[i * k + j] * d[i] * d[j];
    if (!cholesky(L, k)) return false;
    // A <- L^-1 A L^-T
    for (int j = 0; j < k; j++)  // columns: solve L Y = A
        for (int i = 0; i < k; i++) {
            double s = A[i * k + j];
            for (int p = 0; p < i; p++) s -= L[i * k + p] * A[p * k + j];
            A[i * k + j] = s / L[i * k + i];
        }
    for (int i = 0; i < k; i++)  // rows: solve Z L^T = Y
        for (int j = 0; j < k; j++) {
            double s = A[i * k + j];
            for (int p = 0; p < j; p++) s -= A[i * k + p] * L[j * k + p];
            A[i * k + j] = s / L[j * k + j];
        }
    for (int i = 0; i < k; i++)
        for (int j = i + 1; j < k; j++) A[i * k + j] = A[j * k + i] = 0.5 * (A[i * k + j] + A[j * k + i]);
    std::vector<double> V;
    jacobi_eigen(A, k, V);
    std::vector<int> order(k);
    for (int i = 0; i < k; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return A[x * k + x] < A[y * k + y]; });
    C.assign((size_t)k * want, 0.), theta.assign(want, 0.);
    for (int q = 0; q < want; q++) {
        const int col = order[q];
        theta[q] = A[col * k + col];
        // c = D L^-T v
        std::vector<double> y(k);
        for (int i = k - 1; i >= 0; i--) {
            double s = V[i * k + col];
            for (int p = i + 1; p < k; p++) s -= L[p * k + i] * y[p];
            y[i] = s / L[i * k + i];
        }
        for (int i = 0; i < k; i++) C[(size_t)i * want + q] = y[i] * d[i];
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ host: the operators

int apply_S(mof_ctx* ctx, Ops& ops, const double* x, double* y) {
    if (ops.mode == 0) MOF_LAUNCH(k_sell_apply, blocks_for(ops.n, B), B, 0, (int)ops.n, ctx->wSliceBase.p, ctx->wCol.p, ctx->wS.p, x, y);
    else MOF_TRY(vf_apply_operator(ctx, 0., 1., x, y));
    return MOF_OK;
}
int apply_M(mof_ctx* ctx, Ops& ops, const double* x, double* y) {
    if (ops.mode == 0) MOF_LAUNCH(k_sell_apply, blocks_for(ops.n, B), B, 0, (int)ops.n, ctx->wSliceBase.p, ctx->wCol.p, ops.wM.p, x, y);
    else MOF_TRY(vf_apply_operator(ctx, 1., 0., x, y));
    return MOF_OK;
}
int apply_block(mof_ctx* ctx, Ops& ops, bool mass, const double* X, int m, double* Y) {
    for (int j = 0; j < m; j++) MOF_TRY((mass ? apply_M : apply_S)(ctx, ops, X + (size_t)j * ops.n, Y + (size_t)j * ops.n));
    return MOF_OK;
}

// (Re)values the scalar hierarchy for M + K / tau and plans the Chebyshev sharpening of one cycle as the flow solve does (vector_fields.cu):
// lower end of the spectrum of (cycle x system) from a short Lanczos run, the degree that minimises cycles x predicted outer iterations.
int conformal_preconditioner_setup(mof_ctx* ctx, Ops& ops, double tau) {
    ops.tau = tau;
    MOF_TRY(scalar_system_set(ctx, 1. / tau));
    if (!mg_scalar_usable(ctx)) {
        ops.twoCycle = false;
        return MOF_OK;
    }
    double lambdaMin = 0.1;
    MOF_TRY(mg_scalar_smallest_eigenvalue(ctx, 50, &lambdaMin));
    const double rho = std::min(0.9999, std::max(0.3, 1. - lambdaMin));
    ops.chebLo = std::max(1e-4, 0.8 * (1. - rho));
    const double kap = 1.05 / ops.chebLo, q = (std::sqrt(kap) - 1.) / (std::sqrt(kap) + 1.);
    double best = 1e300;
    ops.chebDegree = 2;
    for (int k = 2; k <= 32; k++) {
        const double f = 2. * std::pow(q, k) / (1. + std::pow(q, 2 * k)), cost = k * (1. + f) / (1. - f);
        if (cost < best) best = cost, ops.chebDegree = k;
    }
    ops.twoCycle = true;
    if (getenv("MOF_SPECTRUM_VERBOSE"))
        fprintf(stderr, "[spectrum] two-cycle preconditioner on (K + %.3g M) M^-1 (K + %.3g M) / 2: one cycle contracts by %.3f, Chebyshev degree %d\n", tau, tau, rho, ops.chebDegree);
    return MOF_OK;
}
// W[:, j] = T R[:, j] for the m columns of a block, three columns per application of the six-channel hierarchy.
int conformal_precondition_block(mof_ctx* ctx, Ops& ops, const double* R, int m, double* W) {
    const int V = ctx->V;
    const long long n = ops.n;
    const double eps = 1. / ops.tau, kappa = 2. * eps * eps;
    for (int j0 = 0; j0 < m; j0 += 3) {
        const int cols = std::min(3, m - j0);
        MOF_LAUNCH(k_conformal_pack3, blocks_for(V, B), B, 0, R + (size_t)j0 * n, n, V, cols, ops.r6.p);
        MOF_TRY(mg_scalar_cheb(ctx, ops.r6.p, ops.z6.p, ops.chebDegree, ops.chebLo));
        MOF_LAUNCH(k_conformal_weight6, blocks_for(6ll * V, B), B, 0, (const double*)ops.z6.p, (const double*)ctx->m0.p, V, ops.r6.p);
        MOF_TRY(mg_scalar_cheb(ctx, ops.r6.p, ops.z6.p, ops.chebDegree, ops.chebLo));
        MOF_LAUNCH(k_conformal_unpack3, blocks_for(V, B), B, 0, (const double*)ops.z6.p, kappa, n, V, cols, W + (size_t)j0 * n);
    }
    return MOF_OK;
}

struct Work {
    ScopedBuf<double> partial, gram, coef;
    std::vector<double> host;
};
// G (host, ka x kb, row-major) = A^T B
int gram(mof_ctx* ctx, Work& w, const double* A, const double* Bm, long long n, int ka, int kb, double* G) {
    const int outs = ka * kb;
    const int ctas = (int)std::min<long long>(GRAM_CTAS, (n + GRAM_ROWS - 1) / GRAM_ROWS);
    MOF_LAUNCH(k_gram_partial, ctas, B, 0, A, Bm, n, ka, kb, w.partial.p);
    MOF_LAUNCH(k_gram_fold, blocks_for(outs, B), B, 0, w.partial.p, ctas, outs, w.gram.p);
    MOF_CUDA(read_back(ctx, G, w.gram.p, (size_t)outs));
    return MOF_OK;
}
int combine(mof_ctx* ctx, Work& w, const double* const* in, int count, long long n, int kIn, int kOut, const double* C, double* out) {
    Blocks3 blk;
    blk.count = count;
    for (int b = 0; b < 3; b++) blk.in[b] = b < count ? in[b] : nullptr;
    MOF_CUDA(cudaMemcpyAsync(w.coef.p, C, sizeof(double) * count * kIn * kOut, cudaMemcpyHostToDevice, ctx->stream));
    MOF_LAUNCH(k_combine, blocks_for(n, B), B, 0, blk, n, kIn, kOut, (const double*)w.coef.p, out);
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));  // C is the caller's host memory
    return MOF_OK;
}

}  // namespace

// The `count` lowest eigenpairs of S x = lambda M x for the basis in ctx->params (mesh set, no signals needed). eigenvalues[count]
// ascending; fields[count][T][2] = P x per triangle, x normalised to x^T M x = 1 (ARPACK's normalisation; the sign is free).
int spectrum_lowest(mof_ctx* ctx, int count, double tol, int maxIterations, double* eigenvalues, double* fields, int* iterationsOut, double* residualOut) {
    const int mode = ctx->params.vfMode;
    Ops ops;
    ops.mode = mode;
    ctx->haveSignals = false;  // the data term's buffers are borrowed (D_t = g_t area_t); mof_set_signals restores an alignment
    if (mode != 0) {
        MOF_TRY(vf_init(ctx));
        MOF_TRY(metric_mass_blocks(ctx));
    }
    const long long n = ops.n = vf_unknowns(ctx);
    // Block size: the wanted pairs plus a guard. A pair converges at a rate set by the gap to the first eigenvalue OUTSIDE the block, and a
    // closed surface's eigenvalues come in clusters (a sphere's: 3, 3, 5, 5, 7, 7 ...): a block that ends inside the cluster of the last wanted
    // pair leaves that gap at zero (65 538 vertices, 20 pairs: 545 iterations with a guard of 5, which ends inside the 7 + 7 cluster at 75.4).
    const int guard = std::max(4, count / 2 + 2);
    const int m = std::min<long long>(std::min(MAXM, count + guard), n);
    if (count < 1 || count > m) return fail(ctx, MOF_E_INVALID, "mof_spectrum: between 1 and 28 eigenvectors (and no more than unknowns)");
    MOF_CUDA(ops.tinv.alloc((size_t)n));
    if (mode == 0) {
        MOF_CUDA(ops.wM.alloc(ctx->wS.n));
        MOF_CUDA(cudaMemsetAsync(ops.wM.p, 0, sizeof(double) * ctx->wS.n, ctx->stream));  // (the assembly writes a row's own entries; the slices' padding stays 0)
        MOF_TRY(whitney_mass_operator(ctx, ops.wM.p));
        MOF_LAUNCH(k_sell_inverse_diagonal, blocks_for(n, B), B, 0, (int)n, ctx->wSliceBase.p, ctx->wCol.p, ctx->wS.p, ops.tinv.p);
        const char* e = getenv("MOF_SPECTRUM_MG");
        if (mg_flow_usable(ctx) && !(e && *e == '0')) {
            // The flow hierarchy (multigrid.cu) re-valued for S + tau M. tau > 0 keeps the matrix definite on any genus (harmonic fields are in
            // the null space of S); what it should be is the size of the WANTED eigenvalues — T (S - theta M) then has the spectrum
            // (lambda - theta) / (lambda + tau), clustered at 1 for the bulk and of order one for the wanted pairs — which is not known
            // yet: start from tr S / tr M / 1e3 (far below the bulk whatever the units) and follow the Ritz values down (retune below).
            MOF_CUDA(ctx->dtmp0.reserve((size_t)n));
            MOF_LAUNCH(k_shifted_operator, blocks_for(n, B), B, 0, (int)n, ctx->wSliceBase.p, ctx->wCol.p, ctx->wS.p, ops.wM.p, 0., ctx->wA.p, ctx->wDinv.p);
            MOF_LAUNCH(k_invert_positive, blocks_for(n, B), B, 0, n, ctx->wDinv.p);  // diag S
            MOF_TRY(reduce_sum(ctx, ctx->wDinv.p, n, ctx->scalars.p + SC_TMP));
            MOF_LAUNCH(k_sell_inverse_diagonal, blocks_for(n, B), B, 0, (int)n, ctx->wSliceBase.p, ctx->wCol.p, ops.wM.p, ctx->dtmp0.p);
            MOF_LAUNCH(k_invert_positive, blocks_for(n, B), B, 0, n, ctx->dtmp0.p);  // diag M
            MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, n, ctx->scalars.p + SC_TMP + 1));
            double tr[2] = {0, 0};
            MOF_CUDA(read_back(ctx, tr, ctx->scalars.p + SC_TMP, 2));
            const double tau = ops.tau = tr[1] > 0 ? 1e-3 * tr[0] / tr[1] : 0.;
            MOF_LAUNCH(k_shifted_operator, blocks_for(n, B), B, 0, (int)n, ctx->wSliceBase.p, ctx->wCol.p, ctx->wS.p, ops.wM.p, tau, ctx->wA.p, ctx->wDinv.p);
            ctx->haveFlowSystem = false;
            ops.cycle = mg_flow_try_update(ctx);
            if (getenv("MOF_SPECTRUM_VERBOSE")) fprintf(stderr, "[spectrum] multigrid preconditioner on S + %.3g M: %s\n", tau, ops.cycle ? "yes" : "no (inverse diagonal)");
        }
    } else {
        MOF_TRY(vf_smooth_diagonal(ctx, ops.tinv.p));
        MOF_LAUNCH(k_invert_positive, blocks_for(n, B), B, 0, n, ops.tinv.p);
        const char* e = getenv("MOF_SPECTRUM_MG");
        if (mode == 1 && mg_scalar_usable(ctx) && !(e && *e == '0')) {
            // The inverse diagonal does nothing about the bi-Laplacian's h^-4 conditioning (65 538 vertices: no convergence in 4 000
            // iterations). Start the shift at the scale of the bulk of K's spectrum over 1e3 and follow the Ritz values down (below).
            MOF_CUDA(ops.r6.alloc(6ull * ctx->V));
            MOF_CUDA(ops.z6.alloc(6ull * ctx->V));
            MOF_CUDA(ctx->dtmp0.reserve((size_t)n));
            MOF_TRY(vf_smooth_diagonal(ctx, ctx->dtmp0.p));
            MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, n, ctx->scalars.p + SC_TMP));
            MOF_TRY(reduce_sum(ctx, ctx->m0.p, ctx->V, ctx->scalars.p + SC_TMP + 1));
            double tr[2] = {0, 0};
            MOF_CUDA(read_back(ctx, tr, ctx->scalars.p + SC_TMP, 2));
            // tr S ~ sum_v K_vv^2 / m_v and K_vv ~ lambda_bulk m_v: sqrt(tr S / tr M_lumped) is the bulk of K's generalised spectrum
            const double bulk = tr[1] > 0 ? std::sqrt(std::max(tr[0], 0.) / (2. * tr[1])) : 1.;
            MOF_TRY(conformal_preconditioner_setup(ctx, ops, std::max(1e-3 * bulk, 1e-12)));
        }
    }
    // blocks: X W P and their images under S and M, plus one spare of each for the combinations
    ScopedBuf<double> store;
    const size_t blk = (size_t)n * m;
    MOF_CUDA(store.alloc(12 * blk));
    double *X = store.p, *W = X + blk, *P = W + blk, *SX = P + blk, *SW = SX + blk, *SP = SW + blk, *MX = SP + blk, *MW = MX + blk, *MP = MW + blk, *T0 = MP + blk,
           *T1 = T0 + blk, *T2 = T1 + blk;
    Work w;
    MOF_CUDA(w.partial.alloc((size_t)GRAM_CTAS * MAXM * MAXM));
    MOF_CUDA(w.gram.alloc((size_t)MAXM * MAXM));
    MOF_CUDA(w.coef.alloc((size_t)3 * MAXM * MAXM));
    ScopedBuf<double> dtheta;
    MOF_CUDA(dtheta.alloc(MAXM));

    std::vector<double> theta(m, 0.), G((size_t)m * m), C, th;
    int rc = MOF_OK, it = 0;
    double worst = 0;
    bool haveP = false;
    // Gram blocks between the three parts of the basis: ga[a][b] = part_a^T S part_b, gm likewise (a <= b computed, the rest by symmetry)
    auto body = [&]() -> int {
        MOF_LAUNCH(k_random_block, blocks_for((long long)blk, B), B, 0, n, m, X);
        if (mode == 1) MOF_LAUNCH(k_remove_half_means, 2 * m, B, 0, n / 2, X);
        // Rayleigh-Ritz on the start block
        MOF_TRY(apply_block(ctx, ops, false, X, m, SX));
        MOF_TRY(apply_block(ctx, ops, true, X, m, MX));
        {
            std::vector<double> ga((size_t)m * m), gm((size_t)m * m);
            MOF_TRY(gram(ctx, w, X, SX, n, m, m, ga.data()));
            MOF_TRY(gram(ctx, w, X, MX, n, m, m, gm.data()));
            if (getenv("MOF_SPECTRUM_VERBOSE")) {
                fprintf(stderr, "[spectrum] n %lld m %d; diag(X^T M X):", n, m);
                for (int i = 0; i < m; i++) fprintf(stderr, " %.3g", gm[(size_t)i * m + i]);
                fprintf(stderr, "\n[spectrum] diag(X^T S X):");
                for (int i = 0; i < m; i++) fprintf(stderr, " %.3g", ga[(size_t)i * m + i]);
                fprintf(stderr, "\n[spectrum] row 1 of X^T M X:");
                for (int i = 0; i < m; i++) fprintf(stderr, " %.3g", gm[(size_t)1 * m + i]);
                fprintf(stderr, "\n");
            }
            if (!small_eigen(ga, gm, m, m, C, th)) return fail(ctx, MOF_E_NOCONVERGE, "mof_spectrum: the mass operator is not positive definite on the start block");
            const double* in[1] = {X};
            MOF_TRY(combine(ctx, w, in, 1, n, m, m, C.data(), T0));
            in[0] = SX;
            MOF_TRY(combine(ctx, w, in, 1, n, m, m, C.data(), T1));
            in[0] = MX;
            MOF_TRY(combine(ctx, w, in, 1, n, m, m, C.data(), T2));
            std::swap(X, T0), std::swap(SX, T1), std::swap(MX, T2);
            theta = th;
        }
        for (it = 0; it < maxIterations; it++) {
            if (it % 8 == 7) {  // the images are carried along by the same combinations as the vectors: refresh them before rounding adds up
                MOF_TRY(apply_block(ctx, ops, false, X, m, SX));
                MOF_TRY(apply_block(ctx, ops, true, X, m, MX));
                if (haveP) {
                    MOF_TRY(apply_block(ctx, ops, false, P, m, SP));
                    MOF_TRY(apply_block(ctx, ops, true, P, m, MP));
                }
            }
            // residuals and their norms against ||S x|| + theta ||M x||
            MOF_CUDA(cudaMemcpyAsync(dtheta.p, theta.data(), sizeof(double) * m, cudaMemcpyHostToDevice, ctx->stream));
            MOF_LAUNCH(k_residual_block, blocks_for((long long)blk, B), B, 0, n, m, SX, MX, dtheta.p, ops.tinv.p, T0, W);
            if (mode == 1) MOF_LAUNCH(k_remove_half_means, 2 * m, B, 0, n / 2, W);
            if (ops.cycle)
                for (int j = 0; j < m; j++) MOF_TRY(mg_flow_cycle(ctx, T0 + (size_t)j * n, W + (size_t)j * n));
            if (ops.twoCycle) {
                MOF_TRY(conformal_precondition_block(ctx, ops, T0, m, W));
                MOF_LAUNCH(k_remove_half_means, 2 * m, B, 0, n / 2, W);
            }
            std::vector<double> rr((size_t)m * m), ss((size_t)m * m), mm((size_t)m * m);
            MOF_TRY(gram(ctx, w, T0, T0, n, m, m, rr.data()));
            MOF_TRY(gram(ctx, w, SX, SX, n, m, m, ss.data()));
            MOF_TRY(gram(ctx, w, MX, MX, n, m, m, mm.data()));
            worst = 0;
            // (harmonic fields of a surface of genus > 0 have lambda = 0 and S x -> 0: their residual is measured against the scale of the wanted
            // eigenvalues instead of against itself)
            const double floorTheta = 1e-3 * std::fabs(theta[count - 1]);
            for (int j = 0; j < count; j++) {
                const double den = std::sqrt(ss[j * m + j]) + std::max(std::fabs(theta[j]), floorTheta) * std::sqrt(mm[j * m + j]);
                worst = std::max(worst, den > 0 ? std::sqrt(rr[j * m + j]) / den : 0.);
            }
            if (getenv("MOF_SPECTRUM_VERBOSE") && it % 20 == 0) fprintf(stderr, "[spectrum] iteration %d: residual %.3g, theta[0] %.10g theta[%d] %.10g\n", it, worst, theta[0], count - 1, theta[count - 1]);
            if (!(worst > tol)) break;
            if (!std::isfinite(worst)) return fail(ctx, MOF_E_NOCONVERGE, "mof_spectrum: the iteration broke down");
            // follow the Ritz values down with the preconditioner's shift (see above): re-value the hierarchy when the largest wanted one
            // has fallen below half the shift in use; the directions of the old preconditioner are dropped with it
            if (ops.twoCycle && it >= 3 && theta[count - 1] > 0 && theta[count - 1] < 0.5 * ops.tau) {
                MOF_TRY(conformal_preconditioner_setup(ctx, ops, theta[count - 1]));
                haveP = false;
            }
            if (ops.cycle && it >= 3 && theta[count - 1] > 0 && theta[count - 1] < 0.5 * ops.tau) {
                ops.tau = theta[count - 1];
                MOF_LAUNCH(k_shifted_operator, blocks_for(n, B), B, 0, (int)n, ctx->wSliceBase.p, ctx->wCol.p, ctx->wS.p, ops.wM.p, ops.tau, ctx->wA.p, ctx->wDinv.p);
                ops.cycle = mg_flow_try_update(ctx);
                haveP = false;
                if (getenv("MOF_SPECTRUM_VERBOSE")) fprintf(stderr, "[spectrum] iteration %d: preconditioner re-valued for S + %.3g M (%s)\n", it, ops.tau, ops.cycle ? "cycle" : "inverse diagonal");
            }
            // W against X in the M inner product: W <- W - X (X^T M W)   (X is M-orthonormal)
            MOF_TRY(gram(ctx, w, MX, W, n, m, m, G.data()));
            {
                std::vector<double> Cw((size_t)2 * m * m, 0.);
                for (int i = 0; i < m; i++) Cw[(size_t)i * m + i] = 1;                                      // W
                for (int i = 0; i < m; i++)
                    for (int j = 0; j < m; j++) Cw[(size_t)(m + i) * m + j] = -G[(size_t)i * m + j];        // - X G
                const double* in[2] = {W, X};
                MOF_TRY(combine(ctx, w, in, 2, n, m, m, Cw.data(), T0));
                std::swap(W, T0);
            }
            MOF_TRY(apply_block(ctx, ops, false, W, m, SW));
            MOF_TRY(apply_block(ctx, ops, true, W, m, MW));
            // Rayleigh-Ritz on [X W P]
            const int parts = haveP ? 3 : 2, k = parts * m;
            const double* Z[3] = {X, W, P};
            const double* SZ[3] = {SX, SW, SP};
            const double* MZ[3] = {MX, MW, MP};
            std::vector<double> GA((size_t)k * k), GM((size_t)k * k);
            for (int a = 0; a < parts; a++)
                for (int b = a; b < parts; b++) {
                    MOF_TRY(gram(ctx, w, Z[a], SZ[b], n, m, m, G.data()));
                    for (int i = 0; i < m; i++)
                        for (int j = 0; j < m; j++) GA[(size_t)(a * m + i) * k + b * m + j] = GA[(size_t)(b * m + j) * k + a * m + i] = G[(size_t)i * m + j];
                    MOF_TRY(gram(ctx, w, Z[a], MZ[b], n, m, m, G.data()));
                    for (int i = 0; i < m; i++)
                        for (int j = 0; j < m; j++) GM[(size_t)(a * m + i) * k + b * m + j] = GM[(size_t)(b * m + j) * k + a * m + i] = G[(size_t)i * m + j];
                }
            for (int i = 0; i < k; i++)  // the diagonal blocks are symmetric up to rounding
                for (int j = i + 1; j < k; j++) {
                    GA[(size_t)i * k + j] = GA[(size_t)j * k + i] = 0.5 * (GA[(size_t)i * k + j] + GA[(size_t)j * k + i]);
                    GM[(size_t)i * k + j] = GM[(size_t)j * k + i] = 0.5 * (GM[(size_t)i * k + j] + GM[(size_t)j * k + i]);
                }
            if (!small_eigen(GA, GM, k, m, C, th)) {
                if (haveP) {  // the basis lost its conditioning: drop the directions and go on from [X W]
                    haveP = false;
                    continue;
                }
                return fail(ctx, MOF_E_NOCONVERGE, "mof_spectrum: the Rayleigh-Ritz basis is singular");
            }
            // new directions P = [W P] C_wp and new iterates X = X C_x + P, with their images
            std::vector<double> Cp((size_t)(parts - 1) * m * m);
            for (int i = 0; i < (parts - 1) * m; i++)
                for (int j = 0; j < m; j++) Cp[(size_t)i * m + j] = C[(size_t)(m + i) * m + j];
            std::vector<double> Cx((size_t)2 * m * m, 0.);
            for (int i = 0; i < m; i++)
                for (int j = 0; j < m; j++) Cx[(size_t)i * m + j] = C[(size_t)i * m + j];
            for (int i = 0; i < m; i++) Cx[(size_t)(m + i) * m + i] = 1;
            double** cur[3][3] = {{&X, &W, &P}, {&SX, &SW, &SP}, {&MX, &MW, &MP}};
            for (int f = 0; f < 3; f++) {
                double*& x = *cur[f][0];
                double*& wv = *cur[f][1];
                double*& p = *cur[f][2];
                const double* inP[2] = {wv, p};
                MOF_TRY(combine(ctx, w, inP, parts - 1, n, m, m, Cp.data(), T0));  // new P
                const double* inX[2] = {x, T0};
                MOF_TRY(combine(ctx, w, inX, 2, n, m, m, Cx.data(), T1));          // new X = X C_x + new P
                std::swap(p, T0), std::swap(x, T1);
            }
            haveP = true;
            theta = th;
        }
        return MOF_OK;
    };
    rc = body();
    if (rc == MOF_OK) {
        if (iterationsOut) *iterationsOut = it;
        if (residualOut) *residualOut = worst;
        if (worst > tol) {
            char msg[160];
            snprintf(msg, sizeof(msg), "Unable to Compute Laplacian Spectrum (LOBPCG: residual %g after %d iterations, asked for %g)", worst, it, tol);
            rc = fail(ctx, MOF_E_NOCONVERGE, msg);
        }
    }
    if (rc == MOF_OK) {
        for (int j = 0; j < count; j++) eigenvalues[j] = theta[j];
        // prolonged eigenvectors (VectorLaplacianSpectrum.inl:31-38)
        auto fieldsOut = [&]() -> int {
            for (int j = 0; j < count; j++) {
                if (mode == 0) MOF_TRY(whitney_triangle_field(ctx, X + (size_t)j * n, ctx->tfield.p));
                else MOF_TRY(vf_triangle_field(ctx, X + (size_t)j * n, ctx->tfield.p));
                MOF_CUDA(cudaMemcpyAsync(fields + (size_t)j * 2 * ctx->T, ctx->tfield.p, sizeof(double) * 2 * ctx->T, cudaMemcpyDeviceToHost, ctx->stream));
            }
            MOF_CUDA(cudaStreamSynchronize(ctx->stream));
            MOF_CUDA(cudaMemsetAsync(ctx->tfield.p, 0, sizeof(double) * 2 * ctx->T, ctx->stream));
            return MOF_OK;
        };
        rc = fieldsOut();
    }
    return rc;  // (the blocks and the operators' scratch go back to the pool with their scopes)
}

}  // namespace mof
