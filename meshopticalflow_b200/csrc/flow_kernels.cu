// The per-iteration part of the alignment loop on the GPU.
//
// Signals live as V x 6 (A rgb, B rgb per vertex), the same interleaving the six-right-hand-side
// smoothing solve uses, so one matrix read serves all six channels and a triangle corner is one
// 48-byte fetch. Per-triangle quantities (walk samples, data term) are written once by one thread
// and combined by gathers in a fixed order: no atomics, deterministic.
#include <algorithm>
#include <thread>

#include "mof_internal.cuh"

namespace mof {

constexpr int B = 256;
constexpr int RED_BLOCKS6 = kSMs * 2;

// ------------------------------------------------------------------------------ small vector ops

__global__ void k_axpby(const double* __restrict__ a, const double* __restrict__ b, double w, long long n, double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i] * w;
}

// out6 = A in6 for a V x V CSR (SparseMatrixInterface::Multiply, SparseMatrixInterface.inl:94-109).
__global__ void k_spmv6(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ in, int n,
                        double* __restrict__ out) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int k = rowptr[row]; k < rowptr[row + 1]; k++) {
        double v = val[k];
        const double* src = in + (size_t)col[k] * 6;
#pragma unroll
        for (int j = 0; j < 6; j++) s[j] += v * src[j];
    }
#pragma unroll
    for (int j = 0; j < 6; j++) out[(size_t)row * 6 + j] = s[j];
}

// Six dot products at once: out[j] = sum_v a[v][j] * (w ? w[v] : b[v][j]). Two stages, fixed order.
__global__ void k_dot6_partial(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ w, int n, double* __restrict__ partial) {
    __shared__ double sh[6][B];
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int v = blockIdx.x * B + threadIdx.x; v < n; v += gridDim.x * B) {
        double wv = w ? w[v] : 0.;
#pragma unroll
        for (int j = 0; j < 6; j++) s[j] += a[(size_t)v * 6 + j] * (w ? wv : b[(size_t)v * 6 + j]);
    }
#pragma unroll
    for (int j = 0; j < 6; j++) sh[j][threadIdx.x] = s[j];
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o)
#pragma unroll
            for (int j = 0; j < 6; j++) sh[j][threadIdx.x] += sh[j][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < 6) partial[blockIdx.x * 6 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void k_dot6_final(const double* __restrict__ partial, int np, double* __restrict__ out, int stride) {
    int j = threadIdx.x;
    if (j >= 6) return;
    double s = 0;
    for (int i = 0; i < np; i++) s += partial[i * 6 + j];
    out[j * stride] = s;
}

static int dot6(mof_ctx* ctx, const double* a, const double* b, const double* w, int n, double* out, int stride) {
    MOF_CUDA(ctx->dtmp1.reserve(RED_BLOCKS6 * 6));
    MOF_LAUNCH(k_dot6_partial, RED_BLOCKS6, B, 0, a, b, w, n, ctx->dtmp1.p);
    MOF_LAUNCH(k_dot6_final, 1, 32, 0, ctx->dtmp1.p, RED_BLOCKS6, out, stride);
    return MOF_OK;
}

// Elementwise products for dot / weighted sums through reduce_sum.
__global__ void k_mul(const double* __restrict__ a, const double* __restrict__ b, long long n, double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
}

// ---------------------------------------------------------------- scalar smoothing (a9) and DoG (a5)

int scalar_system_set(mof_ctx* ctx, double eps) {
    MOF_LAUNCH(k_axpby, blocks_for(ctx->nnzS, B), B, 0, ctx->sMass.p, ctx->sStiff.p, eps, ctx->nnzS, ctx->sSys.p);
    MOF_TRY(extract_inverse_diagonal(ctx, ctx->V, ctx->sRowptr.p, ctx->sCol.p, ctx->sSys.p, ctx->sDinv.p));
    if (mg_scalar_usable(ctx)) MOF_TRY(mg_scalar_update(ctx));
    return MOF_OK;
}

// `sameSystem`: the matrix, its inverse diagonal and the multigrid coarse operators are those of the previous call.
static int smooth_solve(mof_ctx* ctx, double weight, const double* in6, double* out6, double tol, bool sameSystem = false) {
    const int V = ctx->V;
    // sM = sMass + weight * sStiffness (OpticalFlow.cpp:355); b = sMass * x (:363); x0 = the signal itself
    if (!sameSystem) {
        MOF_LAUNCH(k_axpby, blocks_for(ctx->nnzS, B), B, 0, ctx->sMass.p, ctx->sStiff.p, weight, ctx->nnzS, ctx->sSys.p);
        MOF_TRY(extract_inverse_diagonal(ctx, V, ctx->sRowptr.p, ctx->sCol.p, ctx->sSys.p, ctx->sDinv.p));
    }
    MOF_CUDA(ctx->rhs6.reserve(6ull * V));
    MOF_LAUNCH(k_spmv6, blocks_for(V, B), B, 0, ctx->sRowptr.p, ctx->sCol.p, ctx->sMass.p, in6, V, ctx->rhs6.p);
    if (out6 != in6) MOF_CUDA(cudaMemcpyAsync(out6, in6, sizeof(double) * 6 * V, cudaMemcpyDeviceToDevice, ctx->stream));
    int iters = 0;
    double relres = 0;
    MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = MOF_OK;
    bool solved = false;
    if (mg_scalar_usable(ctx) && !sameSystem) rc = mg_scalar_update(ctx);
    if (rc == MOF_OK && mg_scalar_usable(ctx)) {
        // multigrid-preconditioned PCG over the six channels at once; a stalled solve falls through to Jacobi-PCG
        int mrc = mg_scalar_solve(ctx, ctx->rhs6.p, out6, tol, std::min(ctx->params.maxCgIterations, 400), &iters, &relres);
        if (mrc == MOF_OK) solved = true;
        else if (mrc != MOF_E_NOCONVERGE) rc = mrc;
    }
    if (rc == MOF_OK && !solved) {
        // a multigrid solve that broke down may have left NaN in out6: the Jacobi solve starts from the signal again
        if (mg_scalar_usable(ctx)) MOF_CUDA(cudaMemcpyAsync(out6, in6, sizeof(double) * 6 * V, cudaMemcpyDeviceToDevice, ctx->stream));
        int jIters = 0;
        rc = pcg_solve_csr6(ctx, V, ctx->sRowptr.p, ctx->sCol.p, ctx->sSys.p, ctx->sDinv.p, ctx->rhs6.p, out6, false, tol, ctx->params.maxCgIterations, &jIters,
                            &relres);
        iters += jIters;
    }
    ctx->stats.smoothCgIterations += iters, ctx->stats.smoothSolves++, ctx->stats.lastSmoothResidual = relres;
    if (rc != MOF_OK) return rc;
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    MOF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.smoothSolveMs += ms;
    return MOF_OK;
}

// ---------------------------------------------------- the next smoothing solve, under the flow solve
//
// The smoothing systems (M + eps_i S) x = M b depend on the iteration's weight and on the input signals, not on the
// flow: the solve of iteration i+1 can run while the flow system of iteration i is being solved. Both solvers spend
// about a third of a PCG iteration in latency-bound launches on the small multigrid levels, which a second stream
// fills. A worker thread drives the scalar solver on its own stream through a VIEW of the context: a shallow copy
// that shares the read-only operators, the smoothing-only buffers (sSys, sDinv, rhs6, the scalar hierarchy) and the
// input signals, and owns what a solve writes besides those (stream, events, counters, Jacobi-PCG work vectors,
// scratch). The owner joins the worker before it uses the result or changes anything the view reads.
// MOF_SMOOTH_AHEAD=0 keeps everything on one stream.
struct SmoothAhead {
    mof_ctx view;
    std::thread worker;
    bool running = false;
    double weight = 0;
    int rc = MOF_OK;
    cudaEvent_t fence = nullptr;  // owner's stream -> view's stream
    DBuf<double> out6, outLo6;    // results; swapped with smoothed6 / smoothedLo6 when consumed
};

static bool smooth_ahead_enabled() {  // read per iteration: tests flip it inside one process
    const char* e = getenv("MOF_SMOOTH_AHEAD");
    if (e && *e) return *e != '0';
    // Under Nsight Compute (its injection variables are in the environment) the second stream's worker thread is left out unless asked
    // for: ncu 2025.2 crashed the process (SIGSEGV inside the injected library) when two threads captured CUDA graphs at the same
    // time (profiles/README.md, r2i). Results are bit-identical either way.
    static const bool profiled = getenv("NV_NSIGHT_INJECTION_PORT_BASE") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR");
    return !profiled;
}

static void smooth_ahead_join(mof_ctx* ctx) {
    SmoothAhead* a = ctx->ahead;
    if (!a || !a->running) return;
    a->worker.join();
    a->running = false;
    mof_stats& s = a->view.stats;
    ctx->stats.kernelLaunches += s.kernelLaunches, ctx->stats.smoothCgIterations += s.smoothCgIterations, ctx->stats.smoothSolves += s.smoothSolves;
    ctx->stats.smoothSolveMs += s.smoothSolveMs;
    if (s.smoothSolves) ctx->stats.lastSmoothResidual = s.lastSmoothResidual;
    memset(&s, 0, sizeof(s));
}

void smooth_ahead_drain(mof_ctx* ctx) {
    smooth_ahead_join(ctx);
    if (ctx->ahead) ctx->ahead->weight = 0, ctx->ahead->rc = MOF_E_INVALID;
}

void smooth_ahead_destroy(mof_ctx* ctx) {
    SmoothAhead* a = ctx->ahead;
    if (!a) return;
    smooth_ahead_join(ctx);
    a->out6.release(), a->outLo6.release();
    // what the view allocated for itself (on its own stream; the shared buffers are the owner's to free)
    cudaStream_t mine = alloc_stream();
    alloc_stream() = a->view.stream;
    a->view.pcg.r.release(), a->view.pcg.d.release(), a->view.pcg.q.release(), a->view.pcg.partial.release(), a->view.pcg.result.release();
    a->view.dtmp0.release(), a->view.dtmp1.release(), a->view.dtmp2.release();
    cudaStreamSynchronize(a->view.stream);
    alloc_stream() = mine;
    if (a->fence) cudaEventDestroy(a->fence);
    if (a->view.ev0) cudaEventDestroy(a->view.ev0);
    if (a->view.ev1) cudaEventDestroy(a->view.ev1);
    if (a->view.stream) cudaStreamDestroy(a->view.stream);
    delete a;
    ctx->ahead = nullptr;
}

// Starts the smoothing solve(s) of `weight` on the second stream. Failing to start is not an error: the caller's
// next iteration then solves on its own stream as before.
static void smooth_ahead_start(mof_ctx* ctx, double weight) {
    if (!smooth_ahead_enabled() || dist_active(ctx) || !(weight != 0)) return;
    const size_t V = (size_t)ctx->V;
    SmoothAhead* a = ctx->ahead;
    if (!a) {
        a = new SmoothAhead();
        bool ok = cudaStreamCreateWithFlags(&a->view.stream, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&a->view.ev0) == cudaSuccess &&
                  cudaEventCreate(&a->view.ev1) == cudaSuccess && cudaEventCreateWithFlags(&a->fence, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            delete a;
            return;
        }
        memset(&a->view.stats, 0, sizeof(a->view.stats));
        ctx->ahead = a;
    }
    smooth_ahead_join(ctx);
    if (a->out6.alloc(6 * V) != cudaSuccess || (ctx->blend && a->outLo6.alloc(6 * V) != cudaSuccess)) return;
    // refresh the view: everything the scalar solver reads, by value (the view never allocates or frees these)
    mof_ctx& v = a->view;
    v.device = ctx->device, v.params = ctx->params, v.pinned = ctx->pinned;
    v.V = ctx->V, v.T = ctx->T, v.E = ctx->E, v.nnzS = ctx->nnzS;
    v.sRowptr = ctx->sRowptr, v.sCol = ctx->sCol, v.sHe = ctx->sHe, v.sMass = ctx->sMass, v.sStiff = ctx->sStiff, v.sSys = ctx->sSys, v.sDinv = ctx->sDinv, v.rhs6 = ctx->rhs6;
    v.sSliceBase = ctx->sSliceBase, v.sColSell = ctx->sColSell, v.sSysSell = ctx->sSysSell, v.sPadded = ctx->sPadded;
    v.mgs = ctx->mgs, v.mg = nullptr, v.dist = nullptr, v.vf = nullptr, v.ahead = nullptr;
    // the view's stream starts after everything queued on the owner's (the signals, the previous smoothing's readers)
    if (cudaEventRecord(a->fence, ctx->stream) != cudaSuccess || cudaStreamWaitEvent(a->view.stream, a->fence, 0) != cudaSuccess) return;
    a->weight = weight, a->rc = MOF_OK, a->running = true;
    const double* in6 = ctx->sig6.p;
    const double* inLo6 = ctx->blend ? ctx->sigLo6.p : nullptr;
    double* out6 = a->out6.p;
    double* outLo6 = a->outLo6.p;
    const double tol = ctx->params.smoothTol;
    a->worker = std::thread([a, in6, inLo6, out6, outLo6, weight, tol] {
        mof_ctx* view = &a->view;
        int rc = cudaSetDevice(view->device) == cudaSuccess ? MOF_OK : MOF_E_CUDA;
        alloc_stream() = view->stream;
        if (rc == MOF_OK) rc = smooth_solve(view, weight, in6, out6, tol, false);
        if (rc == MOF_OK && inLo6) rc = smooth_solve(view, weight, inLo6, outLo6, tol, true);
        if (rc == MOF_OK && cudaStreamSynchronize(view->stream) != cudaSuccess) rc = MOF_E_CUDA;
        a->rc = rc;
    });
}

// The owner's side: true if the solves of `weight` were done ahead; their results are then in smoothed6 (/ smoothedLo6).
static bool smooth_ahead_take(mof_ctx* ctx, double weight) {
    SmoothAhead* a = ctx->ahead;
    if (!a) return false;
    smooth_ahead_join(ctx);
    const bool hit = a->rc == MOF_OK && a->weight == weight && a->out6.p && (!ctx->blend || a->outLo6.p);
    a->weight = 0;
    if (!hit) return false;
    std::swap(ctx->smoothed6, a->out6);
    if (ctx->blend) std::swap(ctx->smoothedLo6, a->outLo6);
    return true;
}

__global__ void k_sub(const double* a, const double* b, long long n, double* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

// OpticalFlow.cpp:848-853: (x - newAvg) * sqrt(oldVar/newVar) + oldAvg, per channel.
// sc: [j*4 + {0: oldAvg, 1: old x.Mx, 2: newAvg, 3: new x.Mx}]
__global__ void k_dog_finish(const double* __restrict__ x, const double* __restrict__ sc, int V, double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6ll * V) return;
    int j = (int)(i % 6);
    double oldAvg = sc[4 * j], oldVar = sc[4 * j + 1] - oldAvg * oldAvg, newAvg = sc[4 * j + 2], newVar = sc[4 * j + 3] - newAvg * newAvg;
    double scale = sqrt(oldVar / newVar);
    out[i] = (x[i] - newAvg) * scale + oldAvg;
}
// The Channels == 6 branch, OpticalFlow.cpp:855: raw * (1 - w) and DoG * w side by side (here: in two V x 6 arrays).
__global__ void k_dog_blend(const double* __restrict__ raw, double w, long long n, double* __restrict__ dog, double* __restrict__ lo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lo[i] = raw[i] * (1. - w), dog[i] *= w;
}

// Difference-of-Gaussians normalisation, OpticalFlow.cpp:822-857 (3-channel branch), all six
// channels at once. getIntegral (FEM.inl:2081-2098) is the dot product with the barycentric
// vertex areas m0: sum_t sum_j x[v_j] * sqrt(det g_t)/6.
// OpticalFlow.cpp:821: log(max(1, x)) * 255 / log(255)
__global__ void k_log_space(const double* __restrict__ in, long long n, double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = log(fmax(1., in[i])) * 255. / log(255.);
}

int dog_preprocess(mof_ctx* ctx) {
    const int V = ctx->V;
    double* sc = ctx->scalars.p + SC_DOG;
    // --log (OpticalFlow.cpp:821) transforms the COMPARISON signals only; the colours that are advected at the end stay raw6
    const double* cmp = ctx->raw6.p;
    if (ctx->params.logSpace) {
        MOF_CUDA(ctx->log6.alloc(6ull * V));
        MOF_LAUNCH(k_log_space, blocks_for(6ll * V, B), B, 0, ctx->raw6.p, 6ll * V, ctx->log6.p);
        cmp = ctx->log6.p;
    }
    MOF_CUDA(ctx->sig6.alloc(6ull * V));
    MOF_CUDA(ctx->smoothed6.alloc(6ull * V));
    MOF_CUDA(ctx->resampled6.alloc(6ull * V));
    MOF_CUDA(ctx->rhs6.alloc(6ull * V));
    ctx->blend = ctx->params.dogWeight > 0 && ctx->params.dogWeight < 1;
    if (ctx->blend) {
        MOF_CUDA(ctx->sigLo6.alloc(6ull * V));
        MOF_CUDA(ctx->smoothedLo6.alloc(6ull * V));
        MOF_CUDA(ctx->resampledLo6.alloc(6ull * V));
    }
    if (!(ctx->params.dogWeight > 0)) {
        MOF_CUDA(cudaMemcpyAsync(ctx->sig6.p, cmp, sizeof(double) * 6 * V, cudaMemcpyDeviceToDevice, ctx->stream));
        return MOF_OK;
    }
    double* x = ctx->smoothed6.p;  // scratch during setup
    double* mb = ctx->resampled6.p;
    MOF_LAUNCH(k_spmv6, blocks_for(V, B), B, 0, ctx->sRowptr.p, ctx->sCol.p, ctx->sMass.p, cmp, V, mb);
    MOF_TRY(dot6(ctx, cmp, nullptr, ctx->m0.p, V, sc + 0, 4));
    MOF_TRY(dot6(ctx, cmp, mb, nullptr, V, sc + 1, 4));
    MOF_TRY(smooth_solve(ctx, ctx->params.dogSmooth, cmp, x, ctx->params.smoothTol));
    MOF_LAUNCH(k_sub, blocks_for(6ll * V, B), B, 0, cmp, x, 6ll * V, x);
    MOF_LAUNCH(k_spmv6, blocks_for(V, B), B, 0, ctx->sRowptr.p, ctx->sCol.p, ctx->sMass.p, x, V, mb);
    MOF_TRY(dot6(ctx, x, nullptr, ctx->m0.p, V, sc + 2, 4));
    MOF_TRY(dot6(ctx, x, mb, nullptr, V, sc + 3, 4));
    MOF_LAUNCH(k_dog_finish, blocks_for(6ll * V, B), B, 0, x, sc, V, ctx->sig6.p);
    if (ctx->blend) MOF_LAUNCH(k_dog_blend, blocks_for(6ll * V, B), B, 0, cmp, ctx->params.dogWeight, 6ll * V, ctx->sig6.p, ctx->sigLo6.p);
    return MOF_OK;
}

// ------------------------------------------------------------------------ triangle walk (a10)

struct WalkMesh {
    const int* opp;
    const double* xlin;
    const double* xcst;
    const double* g;
    const double* vf;
};

__device__ __forceinline__ double metric_dot(const double* g, double ax, double ay, double bx, double by) {
    return ax * (g[0] * bx + g[1] * by) + ay * (g[1] * bx + g[2] * by);
}

// RiemannianMesh::flow, FEM.inl:902-994, same branch order. The direction field is piecewise
// constant per triangle; it is re-read every minStep of arclength, carried across edges by the
// edge transform, and the walk stops when the carried direction opposes the local field.
__device__ void flow_point(const WalkMesh& m, double flowTime, int& tIdx, double& p0, double& p1, double minStep) {
    const double eps = 0.;
    int inEdge = -1, t = tIdx;
    double dir = flowTime < 0 ? -1. : 1.;
    double left = minStep;
    double v0 = m.vf[2 * t] * dir, v1 = m.vf[2 * t + 1] * dir;
    flowTime *= dir;
    for (int count = 0; count < 1000000; count++) {
        if (!(v0 * v0 + v1 * v1)) break;
        double s = 0;
        int idx = -1;
        double c0 = -p1 / v1, c1 = -p0 / v0, c2 = (1. - p0 - p1) / (v1 + v0);
        if (inEdge != 2 && c0 > 0) { double q = p0 + v0 * c0; if (q >= -eps && q <= 1 + eps && c0 > s) idx = 2, s = c0; }
        if (inEdge != 1 && c1 > 0) { double q = p1 + v1 * c1; if (q >= -eps && q <= 1 + eps && c1 > s) idx = 1, s = c1; }
        if (inEdge != 0 && c2 > 0) { double q = p0 + v0 * c2; if (q >= -eps && q <= 1 + eps && c2 > s) idx = 0, s = c2; }
        if (idx == -1) break;
        double gt[3] = {m.g[3 * t], m.g[3 * t + 1], m.g[3 * t + 2]};
        double vv = metric_dot(gt, v0, v1, v0, v1);
        double squareStep = vv * s * s;
        bool refresh = false;
        if (minStep > 0 && squareStep > left * left) s = left / sqrt(vv), refresh = true;
        if (flowTime < s) { p0 += v0 * flowTime, p1 += v1 * flowTime; break; }
        p0 += v0 * s, p1 += v1 * s, flowTime -= s;
        if (refresh) {
            double f0 = m.vf[2 * t], f1 = m.vf[2 * t + 1];
            if (metric_dot(gt, v0, v1, f0, f1) * dir < 0) break;
            v0 = f0 * dir, v1 = f1 * dir;
            left = minStep;
            inEdge = -1;
        } else {
            int h = 3 * t + idx, o = m.opp[h];
            const double* L = m.xlin + 4 * (size_t)h;
            const double* c = m.xcst + 2 * (size_t)h;
            double q0 = L[0] * p0 + L[1] * p1 + c[0], q1 = L[2] * p0 + L[3] * p1 + c[1];
            double w0 = L[0] * v0 + L[1] * v1, w1 = L[2] * v0 + L[3] * v1;
            p0 = q0, p1 = q1, v0 = w0, v1 = w1;
            t = o / 3, inEdge = o - 3 * t;
            left -= sqrt(squareStep);
        }
    }
    tIdx = t;
}

// ResampleSignal's walk + Sample (OpticalFlow.cpp:206-210, 180-186): thread (t, s) flows the
// centroid of triangle t by len[s] and samples signal s (3 channels of in6) where it lands.
__global__ void k_walk_sample(WalkMesh m, const int* __restrict__ tri, const double* __restrict__ in6, int T, double lenA, double lenB,
                              double* __restrict__ tsample6) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * T) return;
    int t = i >> 1, s = i & 1;
    int tt = t;
    double p0 = 1. / 3, p1 = 1. / 3;
    flow_point(m, s ? lenB : lenA, tt, p0, p1, 1e-2);
    int a = tri[3 * tt], b = tri[3 * tt + 1], c = tri[3 * tt + 2];
    double w0 = 1. - p0 - p1;
#pragma unroll
    for (int k = 0; k < 3; k++)
        tsample6[(size_t)t * 6 + 3 * s + k] = in6[(size_t)a * 6 + 3 * s + k] * w0 + in6[(size_t)b * 6 + 3 * s + k] * p0 + in6[(size_t)c * 6 + 3 * s + k] * p1;
}

// ResampleSignal's scatter + normalisation (OpticalFlow.cpp:211, 215) as a gather: the mean of the
// samples of the triangles around each vertex (one outgoing half-edge per incident triangle).
__global__ void k_vertex_gather(const int* __restrict__ rowptr, const int* __restrict__ he, const double* __restrict__ tsample6, int V, double* __restrict__ out6) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= V) return;
    double s[6] = {0, 0, 0, 0, 0, 0};
    int count = 0;
    for (int k = rowptr[a]; k < rowptr[a + 1]; k++) {
        int h = he[k];
        if (h < 0) continue;
        const double* src = tsample6 + (size_t)(h / 3) * 6;
#pragma unroll
        for (int j = 0; j < 6; j++) s[j] += src[j];
        count++;
    }
#pragma unroll
    for (int j = 0; j < 6; j++) out6[(size_t)a * 6 + j] = s[j] / (double)count;
}

int advect_vertices(mof_ctx* ctx, const double* in6, double lenA, double lenB, double* out6) {
    WalkMesh m = {ctx->opp.p, ctx->xlin.p, ctx->xcst.p, ctx->g.p, ctx->tfield.p};
    MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    MOF_LAUNCH(k_walk_sample, blocks_for(2ll * ctx->T, B), B, 0, m, ctx->tri.p, in6, ctx->T, lenA, lenB, ctx->tsample6.p);
    MOF_LAUNCH(k_vertex_gather, blocks_for(ctx->V, B), B, 0, ctx->sRowptr.p, ctx->sHe.p, ctx->tsample6.p, ctx->V, out6);
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    MOF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.advectMs += ms;
    return MOF_OK;
}

// mof_time_kernel(MOF_K_WALK): the walk kernel alone, along the current flow, sampling the comparison signals.
int time_walk_kernel(mof_ctx* ctx, int reps, float* ms) {
    WalkMesh m = {ctx->opp.p, ctx->xlin.p, ctx->xcst.p, ctx->g.p, ctx->tfield.p};
    MOF_CUDA(ctx->tsample6.reserve(6ull * ctx->T));
    for (int i = 0; i < reps + 2; i++) {
        if (i == 2) MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        MOF_LAUNCH(k_walk_sample, blocks_for(2ll * ctx->T, B), B, 0, m, ctx->tri.p, ctx->sig6.p, ctx->T, -0.5, 0.5, ctx->tsample6.p);
    }
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    MOF_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    *ms /= reps;
    return MOF_OK;
}

// ---------------------------------------------------------------------------- data term (a11)

// SetDataTerm, OpticalFlow.cpp:395-421, with the k<2 right-hand side the optimised reference build
// computes. D = sum_c gamma gamma^T area, rhs = sum_c gamma meanDiff area.
// `lo6` (6-channel blend only, else null): the raw half of the signals, channels 0-2 of the reference's Point<Real,6>.
__global__ void k_data_term(const int* __restrict__ tri, const double* __restrict__ area, const double* __restrict__ sig6, const double* __restrict__ lo6, int T,
                            double* __restrict__ D, double* __restrict__ rhs) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double ar = area[t];
    double d00 = 0, d01 = 0, d10 = 0, d11 = 0, r0 = 0, r1 = 0;
    for (int set = lo6 ? 0 : 1; set < 2; set++) {
        const double* sig = set ? sig6 : lo6;
        const double* v0 = sig + (size_t)tri[3 * t] * 6;
        const double* v1 = sig + (size_t)tri[3 * t + 1] * 6;
        const double* v2 = sig + (size_t)tri[3 * t + 2] * 6;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double f0 = (v0[c] + v0[c + 3]) / 2.0, f1 = (v1[c] + v1[c + 3]) / 2.0, f2 = (v2[c] + v2[c + 3]) / 2.0;
            double meanDiff = ((v0[c] - v0[c + 3]) + (v1[c] - v1[c + 3]) + (v2[c] - v2[c + 3])) / 3;
            double ga = f1 - f0, gb = f2 - f0;
            d00 += ga * ga * ar, d01 += ga * gb * ar, d10 += gb * ga * ar, d11 += gb * gb * ar;
            r0 += ga * meanDiff * ar, r1 += gb * meanDiff * ar;
        }
    }
    (void)d10;
    D[3 * t] = d00, D[3 * t + 1] = d01, D[3 * t + 2] = d11;
    rhs[2 * t] = r0, rhs[2 * t + 1] = r1;
}

// ---------------------------------------------------------- flow system assembly and update (a12, a13)

// Row e of R D P and of R rhs (VectorField.h:51-53), written on the fixed pattern of the smooth
// operator (R D P's 5 entries per row are a subset of it); rowSq[e] = sum of squares of the row,
// for the Frobenius normalisation (:57).
__global__ void k_flow_rows(const int* __restrict__ expanded, const int* __restrict__ reduced, const int* __restrict__ opp, const double* __restrict__ P,
                            const double* __restrict__ D, const double* __restrict__ rhs, const int* __restrict__ wRowptr, const int* __restrict__ sliceBase,
                            const int* __restrict__ wCol, int E, double* __restrict__ wA, double* __restrict__ fb, double* __restrict__ rowSq) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int h[2] = {expanded[e], 0};
    h[1] = opp[h[0]];
    int ids[2][3];
    double contrib[2][3];
    double bsum = 0;
    for (int s = 0; s < 2; s++) {
        if (h[s] < 0) { ids[s][0] = ids[s][1] = ids[s][2] = -1; continue; }
        int t = h[s] / 3, ke = h[s] - 3 * t;
        double pe0 = P[6 * t + 2 * ke], pe1 = P[6 * t + 2 * ke + 1];
        double d00 = D[3 * t], d01 = D[3 * t + 1], d11 = D[3 * t + 2];
        // (R D) restricted to triangle t: p_e^T D
        double rd0 = pe0 * d00 + pe1 * d01, rd1 = pe0 * d01 + pe1 * d11;
        for (int q = 0; q < 3; q++) {
            ids[s][q] = reduced[3 * t + q];
            contrib[s][q] = rd0 * P[6 * t + 2 * q] + rd1 * P[6 * t + 2 * q + 1];
        }
        bsum += pe0 * rhs[2 * t] + pe1 * rhs[2 * t + 1];
    }
    double sq = 0;
    const int len = wRowptr[e + 1] - wRowptr[e];
    for (int jj = 0; jj < len; jj++) {
        const size_t k = sell_pos(sliceBase, e, jj);  // sliced layout: coalesced across the 32 rows of a warp
        int f = wCol[k];
        double v = 0;
        for (int s = 0; s < 2; s++)
            for (int q = 0; q < 3; q++)
                if (ids[s][q] == f) v += contrib[s][q];
        wA[k] = v;
        sq += v * v;
    }
    fb[e] = bsum;
    rowSq[e] = sq;
}

// A = s * (R D P) + w * S, b = s * (R rhs), s = 1/||R D P||_F (VectorField.h:57-67); also the
// inverse diagonal for the Jacobi preconditioner.
__global__ void k_flow_finalize(const int* __restrict__ wRowptr, const int* __restrict__ sliceBase, const int* __restrict__ wCol, const double* __restrict__ wS,
                                double weight, int E, double* __restrict__ scalars, double* __restrict__ wA, double* __restrict__ fb, double* __restrict__ dinv) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    double scale = (double)1. / sqrt(scalars[SC_FROB2]);
    if (e == 0) scalars[SC_DATA_SCALE] = scale;
    double diag = 0;
    const int len = wRowptr[e + 1] - wRowptr[e];
    for (int jj = 0; jj < len; jj++) {
        const size_t k = sell_pos(sliceBase, e, jj);
        double v = wA[k] * scale + wS[k] * weight;
        wA[k] = v;
        if (wCol[k] == e) diag += v;
    }
    fb[e] *= scale;
    dinv[e] = 1. / diag;
}

// Per triangle: y = P x restricted to t, y^T D y (for x . Dt x, VectorField.h:91-93).
__global__ void k_step_terms(const int* __restrict__ reduced, const double* __restrict__ P, const double* __restrict__ D, const double* __restrict__ x, int T,
                             double* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double y0 = 0, y1 = 0;
    for (int q = 0; q < 3; q++) {
        double xv = x[reduced[3 * t + q]];
        y0 += P[6 * t + 2 * q] * xv, y1 += P[6 * t + 2 * q + 1] * xv;
    }
    out[t] = y0 * (D[3 * t] * y0 + D[3 * t + 1] * y1) + y1 * (D[3 * t + 1] * y0 + D[3 * t + 2] * y1);
}

// step = (x.b) / (x.Dt x); coeffs += step * x (VectorField.h:93-99).
__global__ void k_update_coeffs(const double* __restrict__ x, const double* __restrict__ scalars, int E, double* __restrict__ coeffs) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    double denom = scalars[SC_STEP_DEN] * scalars[SC_DATA_SCALE], num = scalars[SC_STEP_NUM];
    double step = denom ? num / denom : 0.0;
    if (step) coeffs[e] += x[e] * step;
}

// GetTriangleVectorField, VectorField.h:107-112: tField = P coeffs.
__global__ void k_triangle_field(const int* __restrict__ reduced, const double* __restrict__ P, const double* __restrict__ coeffs, int T, double* __restrict__ tfield) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double y0 = 0, y1 = 0;
    for (int q = 0; q < 3; q++) {
        double c = coeffs[reduced[3 * t + q]];
        y0 += P[6 * t + 2 * q] * c, y1 += P[6 * t + 2 * q + 1] * c;
    }
    tfield[2 * t] = y0, tfield[2 * t + 1] = y1;
}

// UpdateFlow, OpticalFlow.cpp:424-474, followed by VectorField::UpdateOpticalFlow, VectorField.h:46-104.
int update_flow(mof_ctx* ctx, double sWeight, double vfWeight) {
    const int V = ctx->V, T = ctx->T, E = ctx->E;
    // smoothing of both signals, six channels in one solve (:435)
    const double* smoothed = ctx->sig6.p;
    bool doneAhead = false;
    if (sWeight) {
        doneAhead = smooth_ahead_take(ctx, sWeight);  // solved on the second stream during the previous flow solve?
        if (!doneAhead) MOF_TRY(smooth_solve(ctx, sWeight, ctx->sig6.p, ctx->smoothed6.p, ctx->params.smoothTol));
        smoothed = ctx->smoothed6.p;
    } else
        MOF_CUDA(cudaMemcpyAsync(ctx->smoothed6.p, ctx->sig6.p, sizeof(double) * 6 * V, cudaMemcpyDeviceToDevice, ctx->stream));
    // halfway advection of both (:439)
    MOF_TRY(advect_vertices(ctx, smoothed, -0.5, 0.5, ctx->resampled6.p));
    if (ctx->blend) {
        // channels 0-2 of the 6-channel signals: the same system, a second solve; the same walks, a second sampling
        const double* lo = ctx->sigLo6.p;
        if (sWeight) {
            if (!doneAhead) MOF_TRY(smooth_solve(ctx, sWeight, ctx->sigLo6.p, ctx->smoothedLo6.p, ctx->params.smoothTol, true));
            lo = ctx->smoothedLo6.p;
        }
        MOF_TRY(advect_vertices(ctx, lo, -0.5, 0.5, ctx->resampledLo6.p));
    }
    // data term (:470)
    MOF_LAUNCH(k_data_term, blocks_for(T, B), B, 0, ctx->tri.p, ctx->area.p, ctx->resampled6.p, ctx->blend ? ctx->resampledLo6.p : nullptr, T, ctx->dataD.p,
               ctx->dataRhs.p);
    // the next iteration's smoothing (IterativeOptimization's schedule, OpticalFlow.cpp:1041) runs under this iteration's flow solve
    // (not when the flow solve itself borrows the scalar hierarchy: the Conformal basis' preconditioner)
    if (sWeight && ctx->iterationsDone + 1 < ctx->params.iterations && !vf_uses_scalar_hierarchy(ctx)) smooth_ahead_start(ctx, sWeight * ctx->params.sMultiply);
    if (vf_active(ctx)) return vf_update_flow(ctx, vfWeight);  // Conformal / Connection basis (vector_fields.cu)
    // system (VectorField.h:51-67)
    MOF_CUDA(ctx->dtmp0.reserve((size_t)(E > T ? E : T)));
    MOF_LAUNCH(k_flow_rows, blocks_for(E, B), B, 0, ctx->expanded.p, ctx->reduced.p, ctx->opp.p, ctx->P.p, ctx->dataD.p, ctx->dataRhs.p, ctx->wRowptr.p,
               ctx->wSliceBase.p, ctx->wCol.p, E, ctx->wA.p, ctx->fb.p, ctx->dtmp0.p);
    MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, E, ctx->scalars.p + SC_FROB2));
    MOF_LAUNCH(k_flow_finalize, blocks_for(E, B), B, 0, ctx->wRowptr.p, ctx->wSliceBase.p, ctx->wCol.p, ctx->wS.p, vfWeight, E, ctx->scalars.p, ctx->wA.p,
               ctx->fb.p, ctx->wDinv.p);
    ctx->haveFlowSystem = true;
    // solve (:85)
    int iters = 0;
    double relres = 0;
    MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = MOF_OK;
    if (mg_flow_usable(ctx)) rc = mg_flow_update(ctx);  // Galerkin coarse operators of this system; may find it unusable
    if (rc == MOF_OK) {
        bool solved = false;
        if (mg_flow_usable(ctx)) {
            // multigrid-preconditioned PCG; a stalled solve (a damping estimate that was too optimistic) is not an
            // error: the Jacobi-preconditioned kernel below always converges
            int mgIters = 0;
            int mrc = mg_flow_solve(ctx, ctx->params.flowTol, std::min(ctx->params.maxCgIterations, 1000), &mgIters, &relres);
            iters += mgIters;
            if (mrc == MOF_OK) solved = true;
            else if (mrc != MOF_E_NOCONVERGE) rc = mrc;
        }
        if (rc == MOF_OK && !solved) {
            int jIters = 0;
            rc = pcg_solve_sell(ctx, E, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, ctx->wDinv.p, ctx->fb.p, ctx->fx.p, true, ctx->params.flowTol,
                                ctx->params.maxCgIterations, &jIters, &relres);
            iters += jIters;
        }
    }
    ctx->stats.flowCgIterations += iters, ctx->stats.flowSolves++, ctx->stats.lastFlowResidual = relres;
    if (rc != MOF_OK) return rc;
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    MOF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.flowSolveMs += ms;
    // optimal step and update (:91-103)
    MOF_LAUNCH(k_mul, blocks_for(E, B), B, 0, ctx->fx.p, ctx->fb.p, (long long)E, ctx->dtmp0.p);
    MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, E, ctx->scalars.p + SC_STEP_NUM));
    MOF_LAUNCH(k_step_terms, blocks_for(T, B), B, 0, ctx->reduced.p, ctx->P.p, ctx->dataD.p, ctx->fx.p, T, ctx->dtmp0.p);
    MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, T, ctx->scalars.p + SC_STEP_DEN));
    MOF_LAUNCH(k_update_coeffs, blocks_for(E, B), B, 0, ctx->fx.p, ctx->scalars.p, E, ctx->coeffs.p);
    MOF_LAUNCH(k_triangle_field, blocks_for(T, B), B, 0, ctx->reduced.p, ctx->P.p, ctx->coeffs.p, T, ctx->tfield.p);
    return MOF_OK;
}

// The mass operator of the Whitney basis, M = R (g area) P (VectorLaplacianSpectrum.inl:9-19: vfMass, then restriction * vfMass * prolongation),
// on the pattern of S and in its sliced layout: the flow assembly above with D_t = g_t area_t. Overwrites the data term of the alignment.
__global__ void k_metric_mass(const double* __restrict__ g, const double* __restrict__ area, int T, double* __restrict__ D) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const double a = area[t];
    D[3 * t] = g[3 * t] * a, D[3 * t + 1] = g[3 * t + 1] * a, D[3 * t + 2] = g[3 * t + 2] * a;
}
int metric_mass_blocks(mof_ctx* ctx) {
    MOF_LAUNCH(k_metric_mass, blocks_for(ctx->T, B), B, 0, ctx->g.p, ctx->area.p, ctx->T, ctx->dataD.p);
    MOF_CUDA(cudaMemsetAsync(ctx->dataRhs.p, 0, sizeof(double) * 2 * ctx->T, ctx->stream));
    ctx->haveFlowSystem = false;
    return MOF_OK;
}
int whitney_mass_operator(mof_ctx* ctx, double* wM) {
    const int E = ctx->E;
    MOF_TRY(metric_mass_blocks(ctx));
    MOF_CUDA(ctx->dtmp0.reserve((size_t)E));
    MOF_LAUNCH(k_flow_rows, blocks_for(E, B), B, 0, ctx->expanded.p, ctx->reduced.p, ctx->opp.p, ctx->P.p, ctx->dataD.p, ctx->dataRhs.p, ctx->wRowptr.p,
               ctx->wSliceBase.p, ctx->wCol.p, E, wM, ctx->fb.p, ctx->dtmp0.p);
    return MOF_OK;
}
// tField = P coeffs for any coefficient vector of the Whitney basis (GetTriangleVectorField, VectorField.h:107-112).
int whitney_triangle_field(mof_ctx* ctx, const double* coeffs, double* tfield) {
    MOF_LAUNCH(k_triangle_field, blocks_for(ctx->T, B), B, 0, ctx->reduced.p, ctx->P.p, coeffs, ctx->T, tfield);
    return MOF_OK;
}

// ---------------------------------------------------------------------- texel advection (a15)

// Sample, MeshFlow.inl:66-84.
__device__ __forceinline__ void sample_texture(const unsigned char* __restrict__ tex, int W, int H, double u, double v, int bilinear, double* rgb) {
    v = 1 - v;
    u = fmin(1., fmax(0., u)), v = fmin(1., fmax(0., v));
    u *= W - 1, v *= H - 1;
    int x0 = (int)floor(u), y0 = (int)floor(v);
    if (bilinear) {
        double dx = u - x0, dy = v - y0;
        int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
#pragma unroll
        for (int k = 0; k < 3; k++)
            rgb[k] = (double)tex[3 * (W * y0 + x0) + k] * ((1. - dx) * (1. - dy)) + (double)tex[3 * (W * y0 + x1) + k] * (dx * (1. - dy)) +
                     (double)tex[3 * (W * y1 + x1) + k] * (dx * dy) + (double)tex[3 * (W * y1 + x0) + k] * ((1. - dx) * dy);
    } else
        for (int k = 0; k < 3; k++) rgb[k] = (double)tex[3 * (W * y0 + x0) + k];
}

// InputTextureData::flow, OpticalFlow.cpp:501-515; thread (texel, s). Uncovered texels take the
// vertically flipped input (the viewer's initial value, OpticalFlow.cpp:889).
__global__ void k_advect_texels(WalkMesh m, const int* __restrict__ srcT, const double* __restrict__ srcP, const double* __restrict__ triUV,
                                const unsigned char* __restrict__ texA, const unsigned char* __restrict__ texB, int W, int H, double lenA, double lenB, int bilinear,
                                double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * W * H) return;
    int s = i / (W * H), texel = i - s * W * H;
    const unsigned char* tex = s ? texB : texA;
    double* dst = out + (size_t)i * 3;
    int t = srcT[texel];
    if (t == -1) {
        int y = texel / W, x = texel - y * W;
        for (int k = 0; k < 3; k++) dst[k] = (double)tex[3 * ((H - y - 1) * W + x) + k];
        return;
    }
    double p0 = srcP[2 * texel], p1 = srcP[2 * texel + 1];
    flow_point(m, s ? lenB : lenA, t, p0, p1, 1e-2);
    const double* uv = triUV + 6 * (size_t)t;
    double w0 = 1. - p0 - p1;
    double qu = uv[0] * w0 + uv[2] * p0 + uv[4] * p1, qv = uv[1] * w0 + uv[3] * p0 + uv[5] * p1;
    sample_texture(tex, W, H, qu, qv, bilinear, dst);
}

int advect_texels(mof_ctx* ctx, double alpha, int bilinear) {
    WalkMesh m = {ctx->opp.p, ctx->xlin.p, ctx->xcst.p, ctx->g.p, ctx->tfield.p};
    int n = ctx->texW * ctx->texH;
    MOF_CUDA(ctx->texOut.alloc(6ull * n));
    MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    MOF_LAUNCH(k_advect_texels, blocks_for(2ll * n, B), B, 0, m, ctx->srcT.p, ctx->srcP.p, ctx->triUV.p, ctx->tex[0].p, ctx->tex[1].p, ctx->texW, ctx->texH, -alpha,
               1. - alpha, bilinear, ctx->texOut.p);
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    MOF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.advectMs += ms;
    return MOF_OK;
}

// InputTextureData::flow(frames), OpticalFlow.cpp:517-539: thread (texel, s) carries its sample point along the flow in frames - 1
// equal steps (signal 0 backwards, signal 1 forwards; minimum step 1e-2 * frames) and fetches the texture after each; frame 0
// and uncovered texels are the vertically flipped input. out: [2][frames][W*H][3].
__global__ void k_advect_texels_frames(WalkMesh m, const int* __restrict__ srcT, const double* __restrict__ srcP, const double* __restrict__ triUV,
                                       const unsigned char* __restrict__ texA, const unsigned char* __restrict__ texB, int W, int H, int frames, int bilinear,
                                       double* __restrict__ out) {
    const int n = W * H;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    const int s = i / n, texel = i - s * n;
    const unsigned char* tex = s ? texB : texA;
    double* dst = out + ((size_t)s * frames * n + texel) * 3;
    const int y = texel / W, x = texel - y * W;
    double raw[3];
    for (int k = 0; k < 3; k++) raw[k] = (double)tex[3 * ((H - y - 1) * W + x) + k];
    int t = srcT[texel];
    const int covered = t != -1;
    for (int f = 0; f < (covered ? 1 : frames); f++)
        for (int k = 0; k < 3; k++) dst[(size_t)f * n * 3 + k] = raw[k];
    if (!covered) return;
    double p0 = srcP[2 * texel], p1 = srcP[2 * texel + 1];
    const double length = (s ? 1. : -1.) / (frames - 1);
    for (int f = 1; f < frames; f++) {
        flow_point(m, length, t, p0, p1, 1e-2 * frames);
        const double* uv = triUV + 6 * (size_t)t;
        double w0 = 1. - p0 - p1;
        double qu = uv[0] * w0 + uv[2] * p0 + uv[4] * p1, qv = uv[1] * w0 + uv[3] * p0 + uv[5] * p1;
        sample_texture(tex, W, H, qu, qv, bilinear, dst + (size_t)f * n * 3);
    }
}

int advect_texels_frames(mof_ctx* ctx, int frames, int bilinear) {
    WalkMesh m = {ctx->opp.p, ctx->xlin.p, ctx->xcst.p, ctx->g.p, ctx->tfield.p};
    const int n = ctx->texW * ctx->texH;
    MOF_CUDA(ctx->texOut.alloc(6ull * n * frames));
    MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    MOF_LAUNCH(k_advect_texels_frames, blocks_for(2ll * n, B), B, 0, m, ctx->srcT.p, ctx->srcP.p, ctx->triUV.p, ctx->tex[0].p, ctx->tex[1].p, ctx->texW, ctx->texH, frames,
               bilinear, ctx->texOut.p);
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    MOF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.advectMs += ms;
    return MOF_OK;
}

}  // namespace mof
