// The Conformal and Connection vector-field bases (--vfMode 1|2, --cMode 0|1|2):
// same alignment loop, same data term and walks, a different space of flow fields.
//
//   Conformal  (include/Src/Conformal.inl)   2V unknowns [a(0..V-1); b(0..V-1)]: the flow of triangle t is
//              sum_k ginv grad_k a[v_k] + rotGrad_k / sqrt(det g) b[v_k]; smoothness = the bi-Laplacian
//              K diag(1/m) K / 2 on each half (K = scalar stiffness, m = lumped mass).
//   Connection (include/Src/Connection.inl)  2T unknowns [2t+r]: one tangent vector per triangle in its own chart
//              (prolongation = identity); smoothness = sum over edges of l |v_i - L v_ii|^2 in the metric of i,
//              L the linear part of the edge transform, l one of three edge weights.
//
// The reference forms R D P + w S by sparse products and factorises it (VectorField.h:46-104). Here neither
// system matrix is ever formed: A p is applied from its factors (per-triangle 2x2 data blocks through P, the V x V
// stiffness CSR twice for the bi-Laplacian, per-triangle 2x2 transport blocks for the connection), which reads
// less memory than the assembled matrix would (Conformal: a 2-ring pattern of ~19 entries per row, never stored),
// and the system is solved by PCG with a 2x2 block-Jacobi preconditioner (the (a_v, b_v) pair of a vertex, the
// two components of a triangle's vector). All reductions are two-stage on fixed grids: deterministic.
//
// The Conformal system is singular (constants of either potential are in the null space of P and of K); the
// right-hand side R rhs is in its range, so CG from a zero guess converges and the flow P x is unique.
//
// Block Jacobi does nothing about the bi-Laplacian's h^-4 conditioning (8 656 iterations at 16 k vertices, 42 000 at
// 65 k). Where the scalar multigrid hierarchy of the smoothing solves exists, the Conformal PCG is preconditioned by
//     B^-1 = kappa C M C   on each half,   C = ONE multigrid cycle on  M + eps K  (~ (M + eps K)^-1),
// i.e. the inverse of (w/2)(K + delta M) M^-1 (K + delta M), delta = 1/eps: the bi-Laplacian plus a second-order term
// w delta K that stands in for the data term P^T D P (which acts like a weighted Laplacian) when eps is chosen so that
// the two have comparable diagonals. With exact inverses that needs 70-100 iterations whatever the mesh size
// (measured 4 k - 16 k vertices); one cycle in place of each inverse costs 1.5-2x more iterations and no inner solves.
// MOF_CONFORMAL_MG=0, a mesh too small for the hierarchy, or a stalled solve fall back to block Jacobi.
#include <cmath>

#include "mof_internal.cuh"
#include "vf_kernels.cuh"

namespace mof {

using namespace vfk;

// ----------------------------------------------------------------------------------------- host

struct VfState {
    int mode = 0, cMode = 0;
    long long N = 0;
    DBuf<double> connDiag, connOff;  // Connection: [T][3], [T][3][4]
    DBuf<double> minv, blk, w, u;    // Conformal: [V], [V][4], [T][2], [V][2]
    DBuf<double> binv, b, x, r, z, p, q, partial, sc;
    // Conformal, two-cycle preconditioner
    bool mgPrec = false;             // the scalar hierarchy is there and MOF_CONFORMAL_MG != 0
    bool mgNow = false;              // ... and set up for the system being solved
    double stiffTrace = 0, kappa = 1;
    int chebDegree = 3;            // Chebyshev steps around the cycle per approximate inverse of the two-cycle preconditioner, and the
    double chebLo = 0.12;          // lower end of the interval the polynomial is built for (per system, vf_update_flow)
    DBuf<double> r6, z6;             // [V][6]
    void release() {
        DBuf<double>* all[] = {&connDiag, &connOff, &minv, &blk, &w, &u, &binv, &b, &x, &r, &z, &p, &q, &partial, &sc, &r6, &z6};
        for (auto* d : all) d->release();
    }
};

void vf_destroy(mof_ctx* ctx) {
    if (!ctx->vf) return;
    ctx->vf->release();
    delete ctx->vf;
    ctx->vf = nullptr;
}

bool vf_active(const mof_ctx* ctx) { return ctx->vf && ctx->vf->mode != 0; }
bool vf_uses_scalar_hierarchy(const mof_ctx* ctx) { return vf_active(ctx) && ctx->vf->mode == 1 && ctx->vf->mgPrec; }
long long vf_unknowns(const mof_ctx* ctx) { return vf_active(ctx) ? ctx->vf->N : ctx->E; }
const double* vf_rhs(const mof_ctx* ctx) { return ctx->vf->b.p; }
const double* vf_solution(const mof_ctx* ctx) { return ctx->vf->x.p; }

// VectorField::Init for the mode in ctx->params (Conformal.inl:12-82, Connection.inl:22-104); mode 0 tears the state down.
int vf_init(mof_ctx* ctx) {
    const int mode = ctx->params.vfMode;
    if (mode == 0) {
        vf_destroy(ctx);
        return MOF_OK;
    }
    if (!ctx->vf) ctx->vf = new VfState();
    VfState& s = *ctx->vf;
    const int V = ctx->V, T = ctx->T;
    s.mode = mode, s.cMode = ctx->params.cMode;
    s.mgPrec = s.mgNow = false;
    s.N = mode == 1 ? 2ll * V : 2ll * T;
    MOF_CUDA(s.sc.alloc(S_COUNT));
    MOF_CUDA(s.partial.alloc(3 * RED));  // p.q | r.z, r.r
    MOF_CUDA(s.binv.alloc(3ull * (s.N / 2)));
    DBuf<double>* vecs[] = {&s.b, &s.x, &s.r, &s.z, &s.p, &s.q};
    for (auto* v : vecs) MOF_CUDA(v->alloc((size_t)s.N));
    if (mode == 1) {
        MOF_CUDA(s.minv.alloc(V));
        MOF_CUDA(s.blk.alloc(4ull * V));
        MOF_CUDA(s.w.alloc(2ull * T));
        MOF_CUDA(s.u.alloc(2ull * V));
        // lumped mass = sum of sqrt(det)/6 over the fan = the barycentric vertex area m0 (FEM.inl:474, Conformal.inl:29)
        MOF_LAUNCH(k_invert, blocks_for(V, B), B, 0, ctx->m0.p, V, s.minv.p);
        const char* e = getenv("MOF_CONFORMAL_MG");
        s.mgPrec = mg_scalar_usable(ctx) && !(e && *e == '0');
        if (s.mgPrec) {
            MOF_CUDA(s.r6.alloc(6ull * V));
            MOF_CUDA(s.z6.alloc(6ull * V));
            MOF_CUDA(ctx->dtmp0.reserve((size_t)V));
            MOF_LAUNCH(k_stiffness_diagonal, blocks_for(V, B), B, 0, ctx->sRowptr.p, ctx->sHe.p, ctx->sStiff.p, V, ctx->dtmp0.p);
            MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, V, ctx->scalars.p + SC_TMP));
            MOF_CUDA(read_back(ctx, &s.stiffTrace, ctx->scalars.p + SC_TMP));
        }
    } else {
        MOF_CUDA(s.connDiag.alloc(3ull * T));
        MOF_CUDA(s.connOff.alloc(12ull * T));
        MOF_LAUNCH(k_connection_blocks, blocks_for(T, B), B, 0, ctx->g.p, ctx->area.p, ctx->opp.p, ctx->xlin.p, ctx->xcst.p, s.cMode, T, s.connDiag.p, s.connOff.p);
    }
    MOF_CUDA(ctx->coeffs.alloc((size_t)s.N));
    return MOF_OK;
}

// y = A x on the fixed grid, with the partial sums of x . y in s.partial[0..RED).
static int vf_apply(mof_ctx* ctx, double weight, const double* x, double* y) {
    VfState& s = *ctx->vf;
    const int V = ctx->V, T = ctx->T;
    if (s.mode == 1) {
        MOF_LAUNCH(k_conformal_stage1, RED, B, 0, ctx->tri.p, ctx->g.p, ctx->dataD.p, ctx->scalars.p, ctx->sRowptr.p, ctx->sCol.p, ctx->sStiff.p, s.minv.p, x, V, T, s.w.p,
                   s.u.p);
        MOF_LAUNCH(k_conformal_row, RED, B, 0, ctx->sRowptr.p, ctx->sCol.p, ctx->sHe.p, ctx->sStiff.p, ctx->g.p, s.w.p, s.u.p, weight, x, V, y, s.partial.p);
    } else
        MOF_LAUNCH(k_connection_apply, RED, B, 0, ctx->dataD.p, s.connDiag.p, s.connOff.p, ctx->opp.p, ctx->scalars.p, weight, x, T, y, s.partial.p);
    return MOF_OK;
}

// For the Spectrum tool (spectrum.cu): y = (dataScale * P^T D P + weight * S) x with D from ctx->dataD — (0, 1) is the smoothness
// operator S of the basis, (1, 0) with D_t = g_t area_t its mass operator (VectorLaplacianSpectrum.inl:9-19).
__global__ void k_set_data_scale(double v, double* __restrict__ scalars) { scalars[SC_DATA_SCALE] = v; }
int vf_apply_operator(mof_ctx* ctx, double dataScale, double weight, const double* x, double* y) {
    MOF_LAUNCH(k_set_data_scale, 1, 1, 0, dataScale, ctx->scalars.p);
    return vf_apply(ctx, weight, x, y);
}
// diag(S): Connection — the diagonal of each triangle's 2x2 block; Conformal — (K diag(1/m) K / 2)_vv on both halves.
__global__ void k_connection_diagonal(const double* __restrict__ diag, int T, double* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) out[2 * t] = diag[3 * t], out[2 * t + 1] = diag[3 * t + 2];
}
__global__ void k_conformal_diagonal(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ stiff, const double* __restrict__ minv, int V,
                                     double* __restrict__ out) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double d = 0;
    for (int k = rowptr[v]; k < rowptr[v + 1]; k++) d += stiff[k] * stiff[k] * minv[col[k]];
    out[v] = out[v + V] = 0.5 * d;
}
int vf_smooth_diagonal(mof_ctx* ctx, double* out) {
    VfState& s = *ctx->vf;
    if (s.mode == 1) MOF_LAUNCH(k_conformal_diagonal, blocks_for(ctx->V, B), B, 0, ctx->sRowptr.p, ctx->sCol.p, ctx->sStiff.p, s.minv.p, ctx->V, out);
    else MOF_LAUNCH(k_connection_diagonal, blocks_for(ctx->T, B), B, 0, s.connDiag.p, ctx->T, out);
    return MOF_OK;
}
// tField = P coeffs for any coefficient vector of the basis (Conformal.inl:49-64; Connection: the identity).
int vf_triangle_field(mof_ctx* ctx, const double* coeffs, double* tfield) {
    VfState& s = *ctx->vf;
    if (s.mode == 1) MOF_LAUNCH(k_conformal_field, blocks_for(ctx->T, B), B, 0, ctx->tri.p, ctx->g.p, coeffs, ctx->V, ctx->T, tfield);
    else MOF_CUDA(cudaMemcpyAsync(tfield, coeffs, sizeof(double) * 2 * ctx->T, cudaMemcpyDeviceToDevice, ctx->stream));
    return MOF_OK;
}

static int vf_dot(mof_ctx* ctx, const double* a, const double* b, long long n, double* out) {
    VfState& s = *ctx->vf;
    MOF_LAUNCH(k_dot_partial, RED, B, 0, a, b, n, s.partial.p);
    MOF_LAUNCH(k_fold, 1, B, 0, s.partial.p, RED, 1, out, out);
    return MOF_OK;
}

// z = kappa C M C r on both halves (Conformal, s.mgNow), and the partial sums of r.z where k_pcg_direction expects them.
static int vf_two_cycle_preconditioner(mof_ctx* ctx, double* rz) {
    VfState& s = *ctx->vf;
    const int V = ctx->V;
    // each inverse: MOF_CONFORMAL_CHEB (default 3) Chebyshev steps around the cycle — the preconditioner squares the inverse's error, and
    // one unsmoothed-aggregation cycle of a nearly pure Laplacian (eps grows with refinement) is a rough inverse: 875 outer
    // iterations at 65 538 vertices with one cycle per inverse, no finish at 1M (profiles/r1e_modes_1M.txt)
    // (degree and lower bound are chosen per system from the cycle's measured contraction: vf_update_flow)
    MOF_LAUNCH(k_conformal_pack, blocks_for(V, B), B, 0, s.r.p, V, s.r6.p);
    MOF_TRY(mg_scalar_cheb(ctx, s.r6.p, s.z6.p, s.chebDegree, s.chebLo));
    MOF_LAUNCH(k_conformal_weight, blocks_for(V, B), B, 0, s.z6.p, ctx->m0.p, V, s.r6.p);
    MOF_TRY(mg_scalar_cheb(ctx, s.r6.p, s.z6.p, s.chebDegree, s.chebLo));
    MOF_LAUNCH(k_conformal_unpack, blocks_for(V, B), B, 0, s.z6.p, s.kappa, V, s.z.p);
    // The operator annihilates the constants of either potential, and this preconditioner amplifies them by (eps K / M)^2 relative to
    // everything else — 1e10 at 1M vertices, enough for the rounding of the fp32 cycles to take the iteration over (no convergence in
    // 1 800 iterations): keep z in the complement (an orthogonal projector after a symmetric operator: still symmetric, PCG applies).
    MOF_TRY(reduce_sum(ctx, s.z.p, V, ctx->scalars.p + SC_TMP));
    MOF_TRY(reduce_sum(ctx, s.z.p + V, V, ctx->scalars.p + SC_TMP + 1));
    MOF_LAUNCH(k_conformal_remove_means, blocks_for(V, B), B, 0, ctx->scalars.p + SC_TMP, V, s.z.p);
    MOF_LAUNCH(k_dot_partial, RED, B, 0, s.r.p, s.z.p, s.N, rz);
    return MOF_OK;
}

// Block-Jacobi PCG on the matrix-free operator, x0 = 0. Convergence is read back every `kCheck` iterations; the TRUE
// residual b - A x decides, and restarts the recurrence when it has drifted.
static int vf_pcg(mof_ctx* ctx, double weight, double tol, int maxIters, int* itersOut, double* relresOut) {
    VfState& s = *ctx->vf;
    const long long N = s.N, half = N / 2;
    const int split = s.mode == 1 ? 1 : 0;
    // iterations between convergence read-backs (a two-cycle iteration is ~20x a block-Jacobi one); even, so that a batch leaves the
    // ping-pong slot `cur` where it found it and ONE captured graph of a batch serves the whole solve
    const int kCheck = s.mgNow ? 6 : 26;
    double* sc = s.sc.p;
    MOF_CUDA(cudaMemsetAsync(s.x.p, 0, sizeof(double) * N, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(s.r.p, s.b.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, ctx->stream));
    MOF_TRY(vf_dot(ctx, s.b.p, s.b.p, N, sc + S_BB));
    double bb = 0;
    MOF_CUDA(read_back(ctx, &bb, sc + S_BB));
    *itersOut = 0, *relresOut = 0;
    if (!(bb > 0)) return MOF_OK;  // zero right-hand side: x = 0
    int iters = 0;
    double relres = 1, previous = 1;
    double* rzrr = s.partial.p + RED;
    auto iteration = [&](int& cur) -> int {
        MOF_TRY(vf_apply(ctx, weight, s.p.p, s.q.p));
        MOF_LAUNCH(k_pcg_step, RED, B, 0, s.mgNow ? (const double*)nullptr : s.binv.p, sc, S_RZ0 + cur, s.partial.p, RED, s.p.p, s.q.p, half, split, s.x.p, s.r.p,
                   s.z.p, rzrr);
        if (s.mgNow) MOF_TRY(vf_two_cycle_preconditioner(ctx, rzrr));
        MOF_LAUNCH(k_pcg_direction, RED, B, 0, sc, S_RZ0 + (cur ^ 1), S_RZ0 + cur, rzrr, RED, s.z.p, N, s.p.p);
        cur ^= 1;
        return MOF_OK;
    };
    // An iteration is 3 launches (Connection), 4 (Conformal, block Jacobi) or ~80 (Conformal, two-cycle preconditioner): below ~300 k
    // unknowns the host cannot issue them as fast as the GPU retires them. The first batch runs eagerly (it also sizes every scratch
    // buffer), then the same batch is captured once and replayed: one graph launch and one read-back per batch. MOF_VF_GRAPH=0: eager.
    struct BatchGraph {  // destroyed on every way out of the solve
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        void reset() {
            if (exec) cudaGraphExecDestroy(exec);
            if (graph) cudaGraphDestroy(graph);
            exec = nullptr, graph = nullptr;
        }
        ~BatchGraph() { reset(); }
    } batch;
    cudaGraph_t& graph = batch.graph;
    cudaGraphExec_t& exec = batch.exec;
    long long launchesPerBatch = 0;
    bool triedCapture = false;
    {
        const char* e = getenv("MOF_VF_GRAPH");
        if (e && *e == '0') triedCapture = true;
    }
    auto drop_graph = [&]() { batch.reset(); };
    int rcLoop = MOF_OK;
    for (int restart = 0; restart < 8 && rcLoop == MOF_OK; restart++) {
        int cur = 0;
        MOF_LAUNCH(k_pcg_start, RED, B, 0, s.binv.p, s.r.p, half, split, s.z.p, rzrr);  // z = block Jacobi, r.z, r.r
        if (s.mgNow) MOF_TRY(vf_two_cycle_preconditioner(ctx, rzrr));                  // ... replaced
        MOF_LAUNCH(k_pcg_direction, RED, B, 0, sc, S_RZ0 + cur, -1, rzrr, RED, s.z.p, N, s.p.p);
        bool converged = false;
        while (iters < maxIters && !converged) {
            if (exec) {
                cudaError_t ce = cudaGraphLaunch(exec, ctx->stream);
                if (ce != cudaSuccess) { drop_graph(); return cuda_fail(ctx, ce, "cudaGraphLaunch(vf pcg batch)"); }
                ctx->stats.kernelLaunches += launchesPerBatch;
                iters += kCheck;
            } else {
                for (int k = 0; k < kCheck && rcLoop == MOF_OK; k++, iters++) rcLoop = iteration(cur);
                if (rcLoop != MOF_OK) break;
                if (!triedCapture) {
                    triedCapture = true;
                    const long long before = ctx->stats.kernelLaunches;
                    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                        int c2 = cur, crc = MOF_OK;
                        for (int k = 0; k < kCheck && crc == MOF_OK; k++) crc = iteration(c2);
                        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
                        launchesPerBatch = ctx->stats.kernelLaunches - before;
                        if (crc == MOF_OK && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&exec, graph, 0);
                        if (crc != MOF_OK || ce != cudaSuccess || !exec) drop_graph(), cudaGetLastError();
                    }
                    ctx->stats.kernelLaunches = before;
                }
            }
            double rr = 0;
            cudaError_t re = read_back(ctx, &rr, sc + S_RR);
            if (re != cudaSuccess) { drop_graph(); return cuda_fail(ctx, re, "read_back(vf pcg residual)"); }
            if (!(rr == rr)) { drop_graph(); return fail(ctx, MOF_E_NOCONVERGE, "flow PCG (matrix-free) produced a NaN residual"); }
            converged = rr <= tol * tol * bb;
            static const bool verbose = getenv("MOF_VF_VERBOSE") && *getenv("MOF_VF_VERBOSE") != '0';
            if (verbose && (iters % 50 < kCheck || converged)) fprintf(stderr, "[vf pcg] %d iterations, recurrence residual %.3e\n", iters, sqrt(rr / bb));
        }
        if (rcLoop != MOF_OK) break;
        // true residual
        MOF_TRY(vf_apply(ctx, weight, s.x.p, s.q.p));
        MOF_LAUNCH(k_residual, blocks_for(N, B), B, 0, s.b.p, s.q.p, N, s.r.p);
        MOF_TRY(vf_dot(ctx, s.r.p, s.r.p, N, sc + S_RR));
        double rr = 0;
        MOF_CUDA(read_back(ctx, &rr, sc + S_RR));
        relres = sqrt(rr / bb);
        {
            static const bool verbose = getenv("MOF_VF_VERBOSE") && *getenv("MOF_VF_VERBOSE") != '0';
            if (verbose) fprintf(stderr, "[vf pcg] %d iterations, TRUE residual %.3e (restart %d)\n", iters, relres, restart);
        }
        if (relres <= tol || iters >= maxIters) break;
        // The true residual has a floor: the operator itself (K M^-1 K, entries ~ h^-4) is evaluated with a rounding error of ~2e-8 |b| at
        // 1M vertices, and a restart that gains less than a factor 2 has reached it. Below MOF_ACCEPT_RELRES that is an accepted
        // solve, counted in mof_stats.solvesAboveTolerance like a stagnated solve of the other bases.
        if (restart > 0 && relres > 0.5 * previous && relres <= MOF_ACCEPT_RELRES) break;
        previous = relres;
    }
    *itersOut = iters, *relresOut = relres;
    if (!(relres <= tol)) {
        if (iters < maxIters && relres <= MOF_ACCEPT_RELRES) {
            ctx->stats.solvesAboveTolerance++;
            return MOF_OK;
        }
        return fail(ctx, MOF_E_NOCONVERGE, "flow PCG (matrix-free) hit its iteration cap");
    }
    return MOF_OK;
}

// VectorField::UpdateOpticalFlow, VectorField.h:46-104, from the data term in ctx->dataD / ctx->dataRhs.
int vf_update_flow(mof_ctx* ctx, double vfWeight) {
    VfState& s = *ctx->vf;
    const int V = ctx->V, T = ctx->T;
    const long long N = s.N;
    MOF_CUDA(ctx->dtmp0.reserve((size_t)std::max<long long>(N, T)));
    // ||R D P||_F and the scale (:57)
    if (s.mode == 1) {
        MOF_LAUNCH(k_conformal_rows, blocks_for(V, B), B, 0, ctx->sRowptr.p, ctx->sCol.p, ctx->sHe.p, ctx->opp.p, ctx->sStiff.p, s.minv.p, ctx->g.p, ctx->dataD.p, V,
                   ctx->dtmp0.p, s.blk.p);
        MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, V, ctx->scalars.p + SC_FROB2));
    } else {
        MOF_LAUNCH(k_connection_frob, blocks_for(T, B), B, 0, ctx->dataD.p, T, ctx->dtmp0.p);
        MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, T, ctx->scalars.p + SC_FROB2));
    }
    MOF_LAUNCH(k_set_scale, 1, 1, 0, ctx->scalars.p);
    if (s.mode == 1)
        MOF_LAUNCH(k_conformal_finalize, blocks_for(V, B), B, 0, ctx->sRowptr.p, ctx->sHe.p, ctx->g.p, ctx->dataRhs.p, s.blk.p, ctx->scalars.p, vfWeight, V, s.binv.p,
                   s.b.p);
    else
        MOF_LAUNCH(k_connection_finalize, blocks_for(T, B), B, 0, ctx->dataD.p, ctx->dataRhs.p, s.connDiag.p, ctx->scalars.p, vfWeight, T, s.binv.p, s.b.p);
    ctx->haveFlowSystem = true;
    // solve (:85)
    int iters = 0;
    double relres = 0;
    MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    s.mgNow = false;
    if (s.mode == 1 && s.mgPrec && mg_scalar_usable(ctx)) {
        // eps balances the diagonals of w K / eps and of s P^T D P (the factor was chosen by measurement)
        MOF_LAUNCH(k_block_trace, blocks_for(V, B), B, 0, s.blk.p, V, ctx->dtmp0.p);
        MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, V, ctx->scalars.p + SC_TMP));
        double h[2] = {0, 0};
        MOF_CUDA(read_back(ctx, &h[0], ctx->scalars.p + SC_TMP));
        MOF_CUDA(read_back(ctx, &h[1], ctx->scalars.p + SC_DATA_SCALE));
        const double dataTrace = h[0] * h[1];
        if (dataTrace > 0 && std::isfinite(dataTrace) && s.stiffTrace > 0) {
            const double eps = std::min(1e3, std::max(1e-7, 30. * vfWeight * 2. * s.stiffTrace / dataTrace));
            MOF_TRY(scalar_system_set(ctx, eps));
            s.kappa = 2. * eps * eps / vfWeight;
            s.mgNow = mg_scalar_usable(ctx);
            if (s.mgNow) {
                // How sharp an inverse one cycle is on THIS system (eps — the weight of the Laplacian — grows with refinement, and an
                // unsmoothed-aggregation cycle of a nearly pure Laplacian contracts by 0.8-0.95 only), and from it the Chebyshev
                // polynomial around the cycle: the lower spectral bound from a short Lanczos run with a margin (a Ritz value approaches
                // it from above), the degree from the Chebyshev bound. MOF_CONFORMAL_CHEB /
                // MOF_CONFORMAL_CHEB_MIN fix them instead.
                const char* eDeg = getenv("MOF_CONFORMAL_CHEB");
                const char* eLo = getenv("MOF_CONFORMAL_CHEB_MIN");
                double lambdaMin = 0.1;
                if (!(eDeg && *eDeg && eLo && *eLo)) MOF_TRY(mg_scalar_smallest_eigenvalue(ctx, 50, &lambdaMin));
                const double rho = std::min(0.9999, std::max(0.3, 1. - lambdaMin));
                s.chebLo = eLo && *eLo ? atof(eLo) : std::max(1e-4, 0.8 * (1. - rho));
                // the degree that minimises (cycles per inverse) x (outer iterations ~ (1 + f) / (1 - f), f = the polynomial's error bound)
                const double kap = 1.05 / s.chebLo, q = (std::sqrt(kap) - 1.) / (std::sqrt(kap) + 1.);
                int degree = 2;
                double best = 1e300;
                for (int k = 2; k <= 32; k++) {
                    const double f = 2. * std::pow(q, k) / (1. + std::pow(q, 2 * k)), cost = k * (1. + f) / (1. - f);
                    if (cost < best) best = cost, degree = k;
                }
                s.chebDegree = eDeg && *eDeg ? std::max(1, std::min(32, atoi(eDeg))) : degree;
                if (getenv("MOF_MG_VERBOSE") && *getenv("MOF_MG_VERBOSE") != '0')
                    fprintf(stderr, "[conformal] eps %.3g: one cycle contracts by %.3f; Chebyshev degree %d over [%.3f, 1.05]\n", eps, rho, s.chebDegree, s.chebLo);
            }
        }
    }
    int rc = MOF_E_NOCONVERGE, mgIters = 0;
    if (s.mgNow) {
        // a solve that is not done within 600 iterations had too optimistic an interval: widen it (more Chebyshev steps per inverse) and go on
        for (int attempt = 0; attempt < 3 && rc == MOF_E_NOCONVERGE; attempt++) {
            int its = 0;
            rc = vf_pcg(ctx, vfWeight, ctx->params.flowTol, std::min(ctx->params.maxCgIterations, 600), &its, &relres);
            mgIters += its;
            if (rc == MOF_E_NOCONVERGE) s.chebLo *= 0.4, s.chebDegree = std::min(32, s.chebDegree * 3 / 2 + 1);
        }
        s.mgNow = false;
    }
    // no hierarchy — or a stalled solve, which is not an error: block Jacobi always converges
    if (rc == MOF_E_NOCONVERGE) rc = vf_pcg(ctx, vfWeight, ctx->params.flowTol, ctx->params.maxCgIterations, &iters, &relres);
    iters += mgIters;
    ctx->stats.flowCgIterations += iters, ctx->stats.flowSolves++, ctx->stats.lastFlowResidual = relres;
    if (rc != MOF_OK) return rc;
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    MOF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.flowSolveMs += ms;
    // optimal step and update (:91-103)
    MOF_TRY(vf_dot(ctx, s.x.p, s.b.p, N, ctx->scalars.p + SC_STEP_NUM));
    if (s.mode == 1) MOF_LAUNCH(k_conformal_step_terms, blocks_for(T, B), B, 0, ctx->tri.p, ctx->g.p, ctx->dataD.p, s.x.p, V, T, ctx->dtmp0.p);
    else MOF_LAUNCH(k_connection_step_terms, blocks_for(T, B), B, 0, ctx->dataD.p, s.x.p, T, ctx->dtmp0.p);
    MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, T, ctx->scalars.p + SC_STEP_DEN));
    MOF_LAUNCH(k_vf_update_coeffs, blocks_for(N, B), B, 0, s.x.p, ctx->scalars.p, N, ctx->coeffs.p);
    if (s.mode == 1) MOF_LAUNCH(k_conformal_field, blocks_for(T, B), B, 0, ctx->tri.p, ctx->g.p, ctx->coeffs.p, V, T, ctx->tfield.p);
    else MOF_CUDA(cudaMemcpyAsync(ctx->tfield.p, ctx->coeffs.p, sizeof(double) * 2 * T, cudaMemcpyDeviceToDevice, ctx->stream));
    return MOF_OK;
}

}  // namespace mof
