// TEST INFRASTRUCTURE (CPU tier): meshopticalflow_b200/csrc/vector_fields.cu — kernels AND host driver, the very
// source the GPU build compiles — built for the host through emul_cuda_runtime.h, behind two C entry points the
// test calls with numpy arrays: one full VectorField::UpdateOpticalFlow step for a given data term.
#include "emul_cuda_runtime.h"

#include "../../meshopticalflow_b200/csrc/vector_fields.cu"

namespace mof {
// the one function of another translation unit vector_fields.cu calls (setup_kernels.cu: two-stage device reduction)
int reduce_sum(mof_ctx* ctx, const double* in, long long n, double* out) {
    (void)ctx;
    long double s = 0;
    for (long long i = 0; i < n; i++) s += in[i];
    *out = (double)s;
    return MOF_OK;
}
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
                const int* sCol, const int* sHe, const double* sStiff, const double* m0, int vfMode, int cMode, double vfWeight, double tol, int steps,
                const double* D, const double* rhs, double* outB, double* outX, double* outField, double* outScale, double* outCoeffs, long long* itersOut,
                double* relresOut) {
    mof_ctx c;
    mof_ctx* ctx = &c;
    memset(&c.params, 0, sizeof(c.params));
    memset(&c.stats, 0, sizeof(c.stats));
    c.params.vfMode = vfMode, c.params.cMode = cMode, c.params.flowTol = tol, c.params.maxCgIterations = 200000;
    c.V = V, c.T = T, c.E = 3 * T / 2;
    long long nnz = sRowptr[V];
    adopt(c.g, g, 3 * (size_t)T), adopt(c.area, area, T), adopt(c.opp, opp, 3 * (size_t)T), adopt(c.xlin, xlin, 12 * (size_t)T), adopt(c.xcst, xcst, 6 * (size_t)T);
    adopt(c.tri, tri, 3 * (size_t)T), adopt(c.sRowptr, sRowptr, V + 1), adopt(c.sCol, sCol, nnz), adopt(c.sHe, sHe, nnz), adopt(c.sStiff, sStiff, nnz), adopt(c.m0, m0, V);
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
    mof::vf_destroy(ctx);
    return rc;
}

}  // extern "C"
