// TEST INFRASTRUCTURE (CPU tier): meshopticalflow_b200/csrc/vector_fields.cu — kernels AND host driver, the very
// source the GPU build compiles — built for the host through emul_cuda_runtime.h, behind two C entry points the
// test calls with numpy arrays: one full VectorField::UpdateOpticalFlow step for a given data term.
#include "emul_cuda_runtime.h"

#include <vector>

#include "../../meshopticalflow_b200/csrc/vector_fields.cu"

#ifndef EMUL_WITH_FLOW  // linked alone: stand-ins for what vector_fields.cu calls in other translation units, and the entry point
// What vector_fields.cu calls in other translation units, stood in for on the host.
namespace {
bool g_hierarchy = false;            // "the scalar multigrid hierarchy exists"
std::vector<double> g_chol;          // dense Cholesky factor of M + eps K (lower, row-major)
int g_n = 0, g_cycles = 0;
}  // namespace

namespace mof {
// setup_kernels.cu: two-stage device reduction
int reduce_sum(mof_ctx* ctx, const double* in, long long n, double* out) {
    (void)ctx;
    long double s = 0;
    for (long long i = 0; i < n; i++) s += in[i];
    *out = (double)s;
    return MOF_OK;
}
// multigrid.cu / flow_kernels.cu: the scalar hierarchy. Here: sSys = M + eps K factorised densely, and "one cycle" = the exact
// solve rounded to single precision (the cycle computes in fp32) — the preconditioner's structure, not its quality, is
// what the host tier checks.
bool mg_scalar_usable(const mof_ctx*) { return g_hierarchy; }
int scalar_system_set(mof_ctx* ctx, double eps) {
    const int n = ctx->V;
    g_n = n;
    g_chol.assign((size_t)n * n, 0.0);
    for (int r = 0; r < n; r++)
        for (int k = ctx->sRowptr.p[r]; k < ctx->sRowptr.p[r + 1]; k++) g_chol[(size_t)r * n + ctx->sCol.p[k]] = ctx->sMass.p[k] + eps * ctx->sStiff.p[k];
    for (int j = 0; j < n; j++) {
        double d = g_chol[(size_t)j * n + j];
        for (int k = 0; k < j; k++) d -= g_chol[(size_t)j * n + k] * g_chol[(size_t)j * n + k];
        if (!(d > 0)) return MOF_E_INVALID;
        d = sqrt(d);
        g_chol[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double v = g_chol[(size_t)i * n + j];
            for (int k = 0; k < j; k++) v -= g_chol[(size_t)i * n + k] * g_chol[(size_t)j * n + k];
            g_chol[(size_t)i * n + j] = v / d;
        }
    }
    return MOF_OK;
}
int mg_scalar_cycle(mof_ctx* ctx, const double* r6, double* z6) {
    (void)ctx;
    const int n = g_n;
    std::vector<double> y(n);
    g_cycles++;
    for (int c = 0; c < 6; c++) {
        for (int i = 0; i < n; i++) {
            double v = r6[6 * (size_t)i + c];
            for (int k = 0; k < i; k++) v -= g_chol[(size_t)i * n + k] * y[k];
            y[i] = v / g_chol[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; i--) {
            double v = y[i];
            for (int k = i + 1; k < n; k++) v -= g_chol[(size_t)k * n + i] * y[k];
            y[i] = v / g_chol[(size_t)i * n + i];
        }
        for (int i = 0; i < n; i++) z6[6 * (size_t)i + c] = (double)(float)y[i];
    }
    return MOF_OK;
}
// (the harness' "cycle" is an exact inverse: it contracts completely, and a Chebyshev polynomial around it is that inverse again)
int mg_scalar_smallest_eigenvalue(mof_ctx*, int, double* lambdaMin) { *lambdaMin = 1.; return MOF_OK; }
int mg_scalar_cheb(mof_ctx* ctx, const double* r6, double* z6, int, double) { return mg_scalar_cycle(ctx, r6, z6); }
}  // namespace mof

namespace {
template <class T>
void adopt(mof::DBuf<T>& b, const T* host, size_t n) {
    b.alloc(n);
    memcpy(b.p, host, n * sizeof(T));
}
}  // namespace

extern "C" {

// Runs vf_init + `steps` x vf_update_flow on a context filled from host arrays. D / rhs: [steps][T][3] / [steps][T][2] (the data
// terms the caller computed for each step). Outputs per step: b, x [steps][N], tfield [steps][T][2], scale [steps];
// stats: iterations, last relative residual. Returns the library's status code.
int emul_vf_run(int V, int T, const double* g, const double* area, const int* opp, const double* xlin, const double* xcst, const int* tri, const int* sRowptr,
                const int* sCol, const int* sHe, const double* sStiff, const double* sMass, int hierarchy, const double* m0, int vfMode, int cMode, double vfWeight, double tol, int steps,
                const double* D, const double* rhs, double* outB, double* outX, double* outField, double* outScale, double* outCoeffs, long long* itersOut,
                double* relresOut, int* cyclesOut) {
    mof_ctx c;
    mof_ctx* ctx = &c;
    memset(&c.params, 0, sizeof(c.params));
    memset(&c.stats, 0, sizeof(c.stats));
    c.params.vfMode = vfMode, c.params.cMode = cMode, c.params.flowTol = tol, c.params.maxCgIterations = 200000;
    c.V = V, c.T = T, c.E = 3 * T / 2;
    long long nnz = sRowptr[V];
    adopt(c.g, g, 3 * (size_t)T), adopt(c.area, area, T), adopt(c.opp, opp, 3 * (size_t)T), adopt(c.xlin, xlin, 12 * (size_t)T), adopt(c.xcst, xcst, 6 * (size_t)T);
    adopt(c.tri, tri, 3 * (size_t)T), adopt(c.sRowptr, sRowptr, V + 1), adopt(c.sCol, sCol, nnz), adopt(c.sHe, sHe, nnz), adopt(c.sStiff, sStiff, nnz), adopt(c.sMass, sMass, nnz), adopt(c.m0, m0, V);
    c.sSys.alloc(nnz), c.sDinv.alloc(6 * (size_t)V);
    g_hierarchy = hierarchy != 0, g_cycles = 0;
    c.scalars.alloc(mof::SC_COUNT), c.tfield.alloc(2 * (size_t)T), c.dataD.alloc(3 * (size_t)T), c.dataRhs.alloc(2 * (size_t)T), c.coeffs.alloc(c.E);
    int rc = mof::vf_init(ctx);
    if (rc != MOF_OK) return rc;
    long long N = mof::vf_unknowns(ctx);
    memset(c.coeffs.p, 0, sizeof(double) * N);
    for (int s = 0; s < steps && rc == MOF_OK; s++) {
        memcpy(c.dataD.p, D + 3 * (size_t)T * s, sizeof(double) * 3 * T);
        memcpy(c.dataRhs.p, rhs + 2 * (size_t)T * s, sizeof(double) * 2 * T);
        rc = mof::vf_update_flow(ctx, vfWeight);
        if (rc != MOF_OK) break;
        memcpy(outB + N * s, mof::vf_rhs(ctx), sizeof(double) * N);
        memcpy(outX + N * s, mof::vf_solution(ctx), sizeof(double) * N);
        memcpy(outField + 2 * (size_t)T * s, c.tfield.p, sizeof(double) * 2 * T);
        memcpy(outCoeffs + N * s, c.coeffs.p, sizeof(double) * N);
        outScale[s] = c.scalars.p[mof::SC_DATA_SCALE];
    }
    *itersOut = c.stats.flowCgIterations, *relresOut = c.stats.lastFlowResidual;
    if (cyclesOut) *cyclesOut = g_cycles;
    mof::vf_destroy(ctx);
    return rc;
}

}  // extern "C"
#endif  // EMUL_WITH_FLOW
