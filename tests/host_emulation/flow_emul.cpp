// TEST INFRASTRUCTURE (CPU tier): meshopticalflow_b200/csrc/flow_kernels.cu — the per-iteration part of the alignment
// loop: DoG, smoothing right-hand sides, triangle walks, data term, Whitney system assembly, step and update, texel
// advection; kernels AND host driver, the very source the GPU build compiles — built for the host through
// emul_cuda_runtime.h and linked with vector_fields.cu (vf_emul.cpp, -DEMUL_WITH_FLOW). The linear SOLVERS live in
// other translation units (pcg_kernels.cu, multigrid.cu); plain host conjugate-gradient loops stand in for them HERE, so
// this harness checks everything around the solves tap by tap. library_emul.cpp builds the whole library, solvers included.
#include "emul_cuda_runtime.h"

#include <vector>

#include "../../meshopticalflow_b200/csrc/flow_kernels.cu"

namespace {

// Jacobi-preconditioned CG on a generic "apply" (host, double), nrhs interleaved right-hand sides solved independently.
template <class Apply>
int host_pcg(int n, int nrhs, Apply apply, const double* dinv, int dinvStride, const double* b, double* x, bool zeroGuess, double tol, int maxIters, int* itersOut,
             double* relresOut) {
    std::vector<double> r(n), z(n), p(n), q(n), xs(n), bs(n);
    int worst = 0;
    double worstRes = 0;
    for (int c = 0; c < nrhs; c++) {
        for (int i = 0; i < n; i++) bs[i] = b[(size_t)i * nrhs + c], xs[i] = zeroGuess ? 0. : x[(size_t)i * nrhs + c];
        apply(xs.data(), q.data());
        double bb = 0, rz = 0;
        for (int i = 0; i < n; i++) r[i] = bs[i] - q[i], bb += bs[i] * bs[i];
        int it = 0;
        double rr = 0;
        for (int i = 0; i < n; i++) rr += r[i] * r[i];
        if (bb > 0 && rr > tol * tol * bb) {
            for (int i = 0; i < n; i++) z[i] = r[i] * dinv[(size_t)i * dinvStride], p[i] = z[i], rz += r[i] * z[i];
            while (it < maxIters) {
                apply(p.data(), q.data());
                double pq = 0;
                for (int i = 0; i < n; i++) pq += p[i] * q[i];
                double alpha = rz / pq;
                rr = 0;
                for (int i = 0; i < n; i++) xs[i] += alpha * p[i], r[i] -= alpha * q[i], rr += r[i] * r[i];
                it++;
                if (rr <= tol * tol * bb) break;
                double rzn = 0;
                for (int i = 0; i < n; i++) z[i] = r[i] * dinv[(size_t)i * dinvStride], rzn += r[i] * z[i];
                double beta = rzn / rz;
                rz = rzn;
                for (int i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
            }
        }
        for (int i = 0; i < n; i++) x[(size_t)i * nrhs + c] = xs[i];
        worst = std::max(worst, it);
        if (bb > 0) worstRes = std::max(worstRes, sqrt(rr / bb));
    }
    *itersOut = worst, *relresOut = worstRes;
    return worstRes <= tol ? MOF_OK : MOF_E_NOCONVERGE;
}

template <class T>
void adopt(mof::DBuf<T>& b, const T* host, size_t n) {
    b.alloc(n);
    memcpy(b.p, host, n * sizeof(T));
}

}  // namespace

namespace mof {

// setup_kernels.cu
int reduce_sum(mof_ctx* ctx, const double* in, long long n, double* out) {
    (void)ctx;
    long double s = 0;
    for (long long i = 0; i < n; i++) s += in[i];
    *out = (double)s;
    return MOF_OK;
}
// pcg_kernels.cu
int extract_inverse_diagonal(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, double* dinv) {
    (void)ctx;
    for (int r = 0; r < n; r++)
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++)
            if (col[k] == r) dinv[r] = 1. / val[k];
    return MOF_OK;
}
int pcg_solve_csr6(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, const double* dinv, const double* b, double* x, bool zeroGuess, double tol,
                   int maxIters, int* itersOut, double* relresOut) {
    (void)ctx;
    auto apply = [&](const double* in, double* out) {
        for (int r = 0; r < n; r++) {
            double s = 0;
            for (int k = rowptr[r]; k < rowptr[r + 1]; k++) s += val[k] * in[col[k]];
            out[r] = s;
        }
    };
    return host_pcg(n, 6, apply, dinv, 1, b, x, zeroGuess, tol, maxIters, itersOut, relresOut);
}
int pcg_solve_sell(mof_ctx* ctx, int n, const int* sliceBase, const int* col, const double* val, const double* dinv, const double* b, double* x, bool zeroGuess,
                   double tol, int maxIters, int* itersOut, double* relresOut) {
    (void)ctx;
    auto apply = [&](const double* in, double* out) {
        for (int r = 0; r < n; r++) {
            const int width = (sliceBase[(r >> 5) + 1] - sliceBase[r >> 5]) / 32;
            double s = 0;
            for (int j = 0; j < width; j++) {
                size_t k = sell_pos(sliceBase, r, j);
                s += val[k] * in[col[k]];
            }
            out[r] = s;
        }
    };
    return host_pcg(n, 1, apply, dinv, 1, b, x, zeroGuess, tol, maxIters, itersOut, relresOut);
}
// multigrid.cu: no hierarchies on the host
bool mg_flow_usable(const mof_ctx*) { return false; }
int mg_flow_update(mof_ctx*) { return MOF_E_INVALID; }
bool mg_flow_try_update(mof_ctx*) { return false; }
int mg_flow_cycle(mof_ctx*, const double*, double*) { return MOF_E_INVALID; }
int mg_flow_solve(mof_ctx*, double, int, int*, double*) { return MOF_E_INVALID; }
bool mg_scalar_usable(const mof_ctx*) { return false; }
int mg_scalar_update(mof_ctx*) { return MOF_E_INVALID; }
int mg_scalar_solve(mof_ctx*, const double*, double*, double, int, int*, double*) { return MOF_E_INVALID; }
int mg_scalar_cycle(mof_ctx*, const double*, double*) { return MOF_E_INVALID; }
int mg_scalar_smallest_eigenvalue(mof_ctx*, int, double*) { return MOF_E_INVALID; }
int mg_scalar_cheb(mof_ctx*, const double*, double*, int, double) { return MOF_E_INVALID; }
// dist.cu
bool dist_active(const mof_ctx*) { return false; }

}  // namespace mof

extern "C" {

// One whole alignment of the solver loop on host arrays: DoG (dogWeight), `iterations` x UpdateFlow with the reference's
// weight schedule, final advection of the raw colours. Whitney operators come in as the product stores them (sliced
// layout). Per-iteration taps: tfield [it][T][2], data term [it][T][3], Whitney right-hand side and solution [it][E]; final: signals after DoG (sig6, and the raw half
// of the 6-channel blend), advected colours [V][6].
int emul_flow_run(int V, int T, int E, const int* tri, const int* opp, const double* g, const double* area, const double* xlin, const double* xcst, const int* sRowptr,
                  const int* sCol, const int* sHe, const double* sMass, const double* sStiff, const double* m0, const int* reduced, const int* expanded, const double* P,
                  const int* wRowptr, const int* wSliceBase, int wSlices, long long wPadded, const int* wCol, const double* wS, const double* raw6, int iterations,
                  double sSmooth, double sMultiply, double vfSmooth, double dogWeight, double dogSmooth, int vfMode, int cMode, double flowTol, double smoothTol,
                  double* outSig6, double* outSigLo6, double* outField, double* outData, double* outAdvected6, double* outRhs, double* outX, long long* launchesOut) {
    setenv("MOF_SMOOTH_AHEAD", "0", 1);  // the worker thread would share the emulator's thread/block registers
    mof_ctx c;
    mof_ctx* ctx = &c;
    memset(&c.params, 0, sizeof(c.params));
    memset(&c.stats, 0, sizeof(c.stats));
    c.params.iterations = iterations, c.params.sSmooth = sSmooth, c.params.sMultiply = sMultiply, c.params.vfSmooth = vfSmooth, c.params.vMultiply = 1.0;
    c.params.vfSThreshold = 1e-8, c.params.dogWeight = dogWeight, c.params.dogSmooth = dogSmooth, c.params.flowTol = flowTol, c.params.smoothTol = smoothTol;
    c.params.maxCgIterations = 200000, c.params.vfMode = vfMode, c.params.cMode = cMode;
    c.V = V, c.T = T, c.E = E, c.nnzS = sRowptr[V], c.wSlices = wSlices, c.wPadded = wPadded, c.nnzW = wRowptr[E];
    const size_t nnz = (size_t)sRowptr[V];
    adopt(c.tri, tri, 3 * (size_t)T), adopt(c.opp, opp, 3 * (size_t)T), adopt(c.g, g, 3 * (size_t)T), adopt(c.area, area, T), adopt(c.xlin, xlin, 12 * (size_t)T);
    adopt(c.xcst, xcst, 6 * (size_t)T), adopt(c.sRowptr, sRowptr, V + 1), adopt(c.sCol, sCol, nnz), adopt(c.sHe, sHe, nnz), adopt(c.sMass, sMass, nnz);
    adopt(c.sStiff, sStiff, nnz), adopt(c.m0, m0, V), adopt(c.reduced, reduced, 3 * (size_t)T), adopt(c.expanded, expanded, E), adopt(c.P, P, 6 * (size_t)T);
    adopt(c.wRowptr, wRowptr, E + 1), adopt(c.wSliceBase, wSliceBase, wSlices + 1), adopt(c.wCol, wCol, (size_t)wPadded), adopt(c.wS, wS, (size_t)wPadded);
    adopt(c.raw6, raw6, 6 * (size_t)V);
    c.sSys.alloc(nnz), c.sDinv.alloc(6 * (size_t)V), c.wA.alloc((size_t)wPadded), c.wDinv.alloc(E), c.scalars.alloc(mof::SC_COUNT), c.coeffs.alloc(E), c.tfield.alloc(2 * (size_t)T);
    c.fb.alloc(E), c.fx.alloc(E), c.dataD.alloc(3 * (size_t)T), c.dataRhs.alloc(2 * (size_t)T), c.tsample6.alloc(6 * (size_t)T);
    memset(c.wA.p, 0, sizeof(double) * (size_t)wPadded);  // like build_mesh_operators: padding entries stay 0 for good
    int rc = mof::dog_preprocess(ctx);
    if (rc == MOF_OK) rc = mof::vf_init(ctx);
    if (rc != MOF_OK) return rc;
    memset(c.coeffs.p, 0, sizeof(double) * mof::vf_unknowns(ctx));
    memset(c.tfield.p, 0, sizeof(double) * 2 * T);
    memcpy(outSig6, c.sig6.p, sizeof(double) * 6 * V);
    if (c.blend) memcpy(outSigLo6, c.sigLo6.p, sizeof(double) * 6 * V);
    const double vfDefault[3] = {3e-6, 5e-7, 1e4};
    double sw = sSmooth, vw = vfSmooth > 0 ? vfSmooth : vfDefault[vfMode];
    for (int i = 0; i < iterations; i++) {  // mof_iterate
        rc = mof::update_flow(ctx, sw, vw);
        if (rc != MOF_OK) return rc;
        sw *= sMultiply;
        c.iterationsDone++;
        memcpy(outField + 2 * (size_t)T * i, c.tfield.p, sizeof(double) * 2 * T);
        memcpy(outData + 3 * (size_t)T * i, c.dataD.p, sizeof(double) * 3 * T);
        if (vfMode == 0) memcpy(outRhs + (size_t)E * i, c.fb.p, sizeof(double) * E), memcpy(outX + (size_t)E * i, c.fx.p, sizeof(double) * E);
    }
    rc = mof::advect_vertices(ctx, c.raw6.p, -0.5, 0.5, c.resampled6.p);  // mof_advect_vertices(alpha = 0.5)
    if (rc != MOF_OK) return rc;
    memcpy(outAdvected6, c.resampled6.p, sizeof(double) * 6 * V);
    *launchesOut = c.stats.kernelLaunches;
    mof::vf_destroy(ctx);
    return MOF_OK;
}

// The texel variant of the final advection (InputTextureData::flow): k_advect_texels on a given flow field and texel map.
int emul_advect_texels(int T, const int* opp, const double* g, const double* xlin, const double* xcst, const double* tfield, int W, int H, const int* srcT,
                       const double* srcP, const double* triUV, const unsigned char* texA, const unsigned char* texB, double alpha, int bilinear, double* out) {
    mof_ctx c;
    mof_ctx* ctx = &c;
    memset(&c.stats, 0, sizeof(c.stats));
    c.T = T, c.texW = W, c.texH = H;
    const size_t n = (size_t)W * H;
    adopt(c.opp, opp, 3 * (size_t)T), adopt(c.g, g, 3 * (size_t)T), adopt(c.xlin, xlin, 12 * (size_t)T), adopt(c.xcst, xcst, 6 * (size_t)T), adopt(c.tfield, tfield, 2 * (size_t)T);
    adopt(c.srcT, srcT, n), adopt(c.srcP, srcP, 2 * n), adopt(c.triUV, triUV, 6 * (size_t)T), adopt(c.tex[0], texA, 3 * n), adopt(c.tex[1], texB, 3 * n);
    int rc = mof::advect_texels(ctx, alpha, bilinear);
    if (rc == MOF_OK) memcpy(out, c.texOut.p, sizeof(double) * 6 * n);
    return rc;
}

}  // extern "C"
