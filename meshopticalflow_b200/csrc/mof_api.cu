// C ABI of libmof_b200.so (include/mof_b200.h): argument checking, host<->device staging, call order.
#include <cstring>
#include <vector>

#include "mof_internal.cuh"

using namespace mof;

namespace {

// Every C-ABI call that touches the device starts with one of these: selects the context's GPU and makes its stream
// the one device buffers are allocated on and freed in (stream-ordered pool, see DBuf).
struct StreamScope {
    explicit StreamScope(mof_ctx* c) {
        cudaSetDevice(c->device);
        alloc_stream() = c->stream;
    }
};

// Interleave two V x 3 signals into V x 6 (A rgb, B rgb) and back.
__global__ void k_interleave(const double* __restrict__ a, const double* __restrict__ b, int V, double* __restrict__ out6) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3ll * V) return;
    long long v = i / 3, c = i - 3 * v;
    out6[6 * v + c] = a[i], out6[6 * v + 3 + c] = b[i];
}
__global__ void k_deinterleave(const double* __restrict__ in6, int V, double* __restrict__ a, double* __restrict__ b) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3ll * V) return;
    long long v = i / 3, c = i - 3 * v;
    a[i] = in6[6 * v + c], b[i] = in6[6 * v + 3 + c];
}

// Indices are 32-bit like the reference's (SparseMatrix<T, int>): the Whitney pattern holds ~33 entries per vertex, the
// half-edge table 8 slots per triangle. Beyond these sizes an index would wrap (and one B200 could not hold the operators).
int check_mesh_size(mof_ctx* ctx, int V, int T, const char* who) {
    if (V < 3 || T < 1) return fail(ctx, MOF_E_INVALID, std::string(who) + ": empty mesh");
    // ~33 entries per vertex, up to ~20 % slice padding on irregular meshes: 48M vertices keep the padded count below 2^31 (checked again after the scan)
    if (V > 48000000 || T > 96000000) return fail(ctx, MOF_E_INVALID, std::string(who) + ": more than 48M vertices / 96M triangles do not fit 32-bit matrix indices");
    return MOF_OK;
}

// mof_set_reorder, overridden by MOF_REORDER=0|1 in the environment (A/B timing, tests).
int reorder_mode(const mof_ctx* ctx) {
    const char* e = getenv("MOF_REORDER");
    if (e && (*e == '0' || *e == '1')) return *e - '0';
    return ctx->reorderMode;
}
int require_callers_numbering(mof_ctx* ctx, const char* who) {
    if (!ctx->reordered) return MOF_OK;
    return fail(ctx, MOF_E_UNSUPPORTED, std::string(who) + ": the texture map follows the caller's triangle order (the first-writer rule of MeshFlow.inl:281-337); "
                                                            "call mof_set_reorder(ctx, 0) before mof_set_mesh");
}

int require_mesh(mof_ctx* ctx) { return ctx->haveMesh ? MOF_OK : fail(ctx, MOF_E_INVALID, "call mof_set_mesh first"); }
int require_signals(mof_ctx* ctx) { return ctx->haveSignals ? MOF_OK : fail(ctx, MOF_E_INVALID, "call mof_set_signals first"); }

int finish_mesh(mof_ctx* ctx) {
    smooth_ahead_drain(ctx);
    ctx->haveMesh = ctx->haveSignals = ctx->haveFlowSystem = ctx->haveTexture = false;
    cudaEventRecord(ctx->ev0, ctx->stream);
    PhaseTimer pt(ctx);
    int rc = build_mesh_operators(ctx);
    if (rc != MOF_OK) return rc;
    pt.mark("mesh operators (total)");
    rc = mg_setup_mesh(ctx);
    if (rc != MOF_OK) return rc;
    pt.mark("multigrid hierarchies (total)");
    rc = dist_setup_mesh(ctx);
    if (rc != MOF_OK) return rc;
    rc = mg_dist_setup(ctx);
    if (rc != MOF_OK) return rc;
    rc = dist_p2p_setup(ctx);
    if (rc != MOF_OK) return rc;
    pt.mark("row blocks and halo lists");
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    MOF_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.setupMs += ms;
    ctx->stats.flowRows = ctx->E, ctx->stats.flowNnz = ctx->nnzW;
    ctx->stats.flowSpmvBytes = 12. * (double)ctx->nnzW + 4. * ((double)ctx->E + 1) + 16. * (double)ctx->E;
    ctx->haveMesh = true;
    return MOF_OK;
}

int finish_signals(mof_ctx* ctx) {
    ctx->haveSignals = false;
    smooth_ahead_drain(ctx);
    if (ctx->params.vfMode != 0 && dist_active(ctx))
        return fail(ctx, MOF_E_UNSUPPORTED, "[ERROR] a partitioned mesh (mof_dist_init) supports the Whitney vector field only");
    mg_new_pair(ctx);  // same inputs -> same bits: nothing carried over from the previous pair's systems
    MOF_TRY(dog_preprocess(ctx));
    MOF_TRY(vf_init(ctx));  // VectorField::Init for --vfMode 1|2 (OpticalFlow.cpp:862-871); Whitney was built with the mesh
    MOF_CUDA(cudaMemsetAsync(ctx->coeffs.p, 0, sizeof(double) * vf_unknowns(ctx), ctx->stream));
    MOF_CUDA(cudaMemsetAsync(ctx->tfield.p, 0, sizeof(double) * 2 * ctx->T, ctx->stream));
    ctx->iterationsDone = 0;
    ctx->curSmooth = ctx->params.sSmooth;
    // _main, OpticalFlow.cpp:1064-1069: the smoothing weight's default depends on the basis
    const double vfDefault[3] = {3e-6, 5e-7, 1e4};
    ctx->curVf = ctx->params.vfSmooth > 0 ? ctx->params.vfSmooth : vfDefault[ctx->params.vfMode];
    ctx->haveSignals = true, ctx->haveFlowSystem = false;
    return MOF_OK;
}

}  // namespace

extern "C" {

void mof_default_params(mof_params* p) {
    if (!p) return;
    p->iterations = 10;
    p->sSmooth = (double)3e-3f;       // cmdLineParameter<float>, OpticalFlow.cpp:59, cast to Real at :1062
    p->sMultiply = (double)0.25f;
    p->vfSmooth = 3e-6;               // double literal, OpticalFlow.cpp:1067
    p->vMultiply = (double)1.0f;
    p->vfSThreshold = (double)1e-8f;
    p->dogWeight = (double)1.f;
    p->dogSmooth = (double)(float)1e-4;
    p->flowTol = 1e-8;
    p->smoothTol = 1e-10;
    p->maxCgIterations = 100000;
    p->vfMode = 0;                    // WHITNEY_VECTOR_FIELD, OpticalFlow.cpp:58
    p->cMode = 0;                     // PROJECTED_BARICENTRIC_WEIGHTS
    p->logSpace = 0;
}

int mof_create(int device, void* stream, mof_ctx** out) {
    if (!out) return MOF_E_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) return MOF_E_CUDA;  // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return MOF_E_CUDA;
    mof_ctx* ctx = new mof_ctx();
    ctx->device = device;
    mof_default_params(&ctx->params);
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    {
        const char* pd = getenv("MOF_PDL");  // programmatic dependent launches in the solvers (mof_internal.cuh); MOF_PDL=0: plain stream order
        ctx->pdl = !(pd && *pd == '0');
    }
    if (stream) ctx->stream = (cudaStream_t)stream;
    else {
        // The context's own stream carries the critical path (the flow solves); the smoothing solves that run ahead on a
        // second stream (flow_kernels.cu) stay at the default, lowest priority and fill what it leaves free.
        int least = 0, greatest = 0;
        const char* pe = getenv("MOF_STREAM_PRIORITY");
        const bool prioritise = !(pe && *pe == '0') && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess && greatest < least;
        cudaError_t se = prioritise ? cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, greatest) : cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) { delete ctx; return MOF_E_CUDA; }
        ctx->ownStream = true;
    }
    if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) { delete ctx; return MOF_E_CUDA; }
    if (cudaHostAlloc((void**)&ctx->pinned, 256 * sizeof(double), cudaHostAllocDefault) != cudaSuccess) { delete ctx; return MOF_E_CUDA; }
    // keep freed blocks in the stream-ordered pool instead of handing them back to the driver at every synchronisation
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = ctx;
    return MOF_OK;
}

void mof_destroy(mof_ctx* ctx) {
    if (!ctx) return;
    StreamScope scope(ctx);
    smooth_ahead_destroy(ctx);
    cudaStreamSynchronize(ctx->stream);
    DBuf<double>* dbl[] = {&ctx->pos, &ctx->g, &ctx->area, &ctx->xlin, &ctx->xcst, &ctx->sMass, &ctx->sStiff, &ctx->sSys, &ctx->sSysSell, &ctx->sDinv, &ctx->P, &ctx->m0, &ctx->m1,
                           &ctx->wS, &ctx->wA, &ctx->wDinv, &ctx->raw6, &ctx->log6, &ctx->sig6, &ctx->smoothed6, &ctx->rhs6, &ctx->resampled6, &ctx->tsample6, &ctx->dataD,
                           &ctx->dataRhs, &ctx->coeffs, &ctx->tfield, &ctx->fb, &ctx->fx, &ctx->scalars, &ctx->pcg.r, &ctx->pcg.d, &ctx->pcg.q, &ctx->pcg.partial,
                           &ctx->pcg.result, &ctx->dtmp0, &ctx->dtmp1, &ctx->dtmp2, &ctx->srcP, &ctx->triUV, &ctx->texOut, &ctx->sigLo6, &ctx->smoothedLo6, &ctx->resampledLo6};
    for (auto* b : dbl) b->release();
    DBuf<int>* ints[] = {&ctx->tri, &ctx->opp, &ctx->sRowptr, &ctx->sCol, &ctx->sSliceBase, &ctx->sColSell, &ctx->sHe, &ctx->reduced, &ctx->expanded, &ctx->positive, &ctx->wRowptr, &ctx->wSliceBase, &ctx->wCol,
                         &ctx->itmp0, &ctx->itmp1, &ctx->itmp2, &ctx->flags, &ctx->srcT};
    for (auto* b : ints) b->release();
    ctx->hashKeys.release(), ctx->tex[0].release(), ctx->tex[1].release();
    ctx->subXyz.release(), ctx->subTri.release(), ctx->subUv.release();
    ctx->vOrder.release(), ctx->vRank.release(), ctx->tOrder.release();
    mg_destroy(ctx);
    dist_destroy(ctx);
    vf_destroy(ctx);
    cudaStreamSynchronize(ctx->stream);  // the frees above are stream-ordered
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->ownStream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* mof_last_error(const mof_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mof_set_params(mof_ctx* ctx, const mof_params* p) {
    if (!ctx || !p) return MOF_E_INVALID;
    smooth_ahead_drain(ctx);
    if (p->vfMode < 0 || p->vfMode > 2) return fail(ctx, MOF_E_INVALID, "ERROR: Unsupported vector field! ");  // OpticalFlow.cpp:867
    if (p->vfMode == 2 && (p->cMode < 0 || p->cMode > 2)) return fail(ctx, MOF_E_INVALID, "Undefined Connection Mode ");  // Connection.inl:68
    if ((p->vfMode != ctx->params.vfMode || p->cMode != ctx->params.cMode || (p->dogWeight != ctx->params.dogWeight)) && ctx->haveSignals)
        ctx->haveSignals = false;  // the basis and the DoG blend are fixed by mof_set_signals: it has to be called again
    if (p->iterations < 0 || !(p->flowTol > 0) || !(p->smoothTol > 0) || p->maxCgIterations < 1) return fail(ctx, MOF_E_INVALID, "bad solver parameters");
    ctx->params = *p;
    return MOF_OK;
}

int mof_get_stats(mof_ctx* ctx, mof_stats* out) {
    if (!ctx || !out) return MOF_E_INVALID;
    *out = ctx->stats;
    return MOF_OK;
}
void mof_reset_stats(mof_ctx* ctx) {
    if (!ctx) return;
    long long rows = ctx->stats.flowRows, nnz = ctx->stats.flowNnz, halo = ctx->stats.haloEntries;
    double bytes = ctx->stats.flowSpmvBytes;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->stats.flowRows = rows, ctx->stats.flowNnz = nnz, ctx->stats.flowSpmvBytes = bytes, ctx->stats.haloEntries = halo;
}
int mof_synchronize(mof_ctx* ctx) {
    if (!ctx) return MOF_E_INVALID;
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

int mof_set_reorder(mof_ctx* ctx, int mode) {
    if (!ctx || mode < -1 || mode > 1) return MOF_E_INVALID;
    ctx->reorderMode = mode;
    return MOF_OK;
}
int mof_get_permutation(mof_ctx* ctx, int* reordered, int* vertexOrder, int* triangleOrder) {
    if (!ctx) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    if (reordered) *reordered = ctx->reordered ? 1 : 0;
    if (!ctx->reordered) {
        if (vertexOrder) for (int i = 0; i < ctx->V; i++) vertexOrder[i] = i;
        if (triangleOrder) for (int i = 0; i < ctx->T; i++) triangleOrder[i] = i;
        return MOF_OK;
    }
    if (vertexOrder) MOF_CUDA(read_back(ctx, vertexOrder, ctx->vOrder.p, (size_t)ctx->V));
    if (triangleOrder) MOF_CUDA(read_back(ctx, triangleOrder, ctx->tOrder.p, (size_t)ctx->T));
    return MOF_OK;
}

int mof_spectrum(mof_ctx* ctx, int count, double tol, int maxIterations, double* eigenvalues, double* fields, int* iterations, double* residual) {
    if (!ctx || !eigenvalues || !fields) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    if (dist_active(ctx)) return fail(ctx, MOF_E_UNSUPPORTED, "mof_spectrum: not on a partitioned mesh");
    if (!(tol > 0) || maxIterations < 1) return fail(ctx, MOF_E_INVALID, "mof_spectrum: bad tolerance / iteration limit");
    StreamScope scope(ctx);
    smooth_ahead_drain(ctx);
    const int T = ctx->T;
    int rc = spectrum_lowest(ctx, count, tol, maxIterations, eigenvalues, fields, iterations, residual);
    if (rc == MOF_OK && ctx->reordered) {  // rows of every field back into the caller's triangle order
        std::vector<double> tmp(2 * (size_t)T);
        std::vector<int> order((size_t)T);
        MOF_CUDA(read_back(ctx, order.data(), ctx->tOrder.p, (size_t)T));
        for (int j = 0; j < count; j++) {
            double* f = fields + 2 * (size_t)T * j;
            for (int t = 0; t < T; t++) tmp[2 * (size_t)order[t]] = f[2 * (size_t)t], tmp[2 * (size_t)order[t] + 1] = f[2 * (size_t)t + 1];
            memcpy(f, tmp.data(), sizeof(double) * 2 * T);
        }
    }
    return rc;
}

int mof_dist_unique_id(unsigned char id128[128]) { return id128 ? dist_unique_id(id128) : MOF_E_INVALID; }
int mof_dist_init(mof_ctx* ctx, int world, int rank, const unsigned char id128[128]) {
    if (!ctx) return MOF_E_INVALID;
    StreamScope scope(ctx);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, MOF_E_CUDA, "cudaSetDevice");
    smooth_ahead_drain(ctx);
    ctx->haveMesh = ctx->haveSignals = false;  // the partition is built with the mesh
    return dist_init(ctx, world, rank, id128);
}

int mof_set_mesh(mof_ctx* ctx, const double* xyz, int V, const int* tri, int T) {
    if (!ctx) return MOF_E_INVALID;
    if (!xyz || !tri) return fail(ctx, MOF_E_INVALID, "mof_set_mesh: empty mesh");
    MOF_TRY(check_mesh_size(ctx, V, T, "mof_set_mesh"));
    StreamScope scope(ctx);
    ctx->V = V, ctx->T = T;
    MOF_CUDA(ctx->pos.alloc(3ull * V));
    MOF_CUDA(ctx->tri.alloc(3ull * T));
    MOF_CUDA(cudaMemcpyAsync(ctx->pos.p, xyz, sizeof(double) * 3 * V, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->tri.p, tri, sizeof(int) * 3 * T, cudaMemcpyHostToDevice, ctx->stream));
    MOF_TRY(reorder_mesh(ctx, reorder_mode(ctx)));
    return finish_mesh(ctx);
}

int mof_set_mesh_device(mof_ctx* ctx, const double* d_xyz, int V, const int* d_tri, int T) {
    if (!ctx) return MOF_E_INVALID;
    if (!d_xyz || !d_tri) return fail(ctx, MOF_E_INVALID, "mof_set_mesh_device: empty mesh");
    MOF_TRY(check_mesh_size(ctx, V, T, "mof_set_mesh_device"));
    StreamScope scope(ctx);
    ctx->V = V, ctx->T = T;
    MOF_CUDA(ctx->pos.alloc(3ull * V));
    MOF_CUDA(ctx->tri.alloc(3ull * T));
    MOF_CUDA(cudaMemcpyAsync(ctx->pos.p, d_xyz, sizeof(double) * 3 * V, cudaMemcpyDeviceToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->tri.p, d_tri, sizeof(int) * 3 * T, cudaMemcpyDeviceToDevice, ctx->stream));
    MOF_TRY(reorder_mesh(ctx, reorder_mode(ctx)));
    return finish_mesh(ctx);
}

static int set_signals_common(mof_ctx* ctx, const double* a, const double* b, int channels, cudaMemcpyKind kind) {
    if (!ctx) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    if (!a || !b) return fail(ctx, MOF_E_INVALID, "mof_set_signals: null signal");
    if (channels != 3) return fail(ctx, MOF_E_UNSUPPORTED, "[ERROR] only 3-channel signals are on the accelerated path (_main<double,3>, OpticalFlow.cpp:1115)");
    StreamScope scope(ctx);
    const int V = ctx->V;
    MOF_CUDA(ctx->raw6.alloc(6ull * V));
    MOF_CUDA(ctx->dtmp0.reserve(6ull * V));
    double* sa = ctx->dtmp0.p;
    double* sb = ctx->dtmp0.p + 3ull * V;
    MOF_CUDA(cudaMemcpyAsync(sa, a, sizeof(double) * 3 * V, kind, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(sb, b, sizeof(double) * 3 * V, kind, ctx->stream));
    if (ctx->reordered) {  // the caller's rows -> the library's numbering (reorder.cu)
        MOF_CUDA(ctx->dtmp2.reserve(6ull * V));
        MOF_TRY(reorder_gather(ctx, 0, sa, 3, ctx->dtmp2.p));
        MOF_TRY(reorder_gather(ctx, 0, sb, 3, ctx->dtmp2.p + 3ull * V));
        sa = ctx->dtmp2.p, sb = ctx->dtmp2.p + 3ull * V;
    }
    MOF_LAUNCH(k_interleave, blocks_for(3ll * V, 256), 256, 0, sa, sb, V, ctx->raw6.p);
    return finish_signals(ctx);
}
int mof_set_signals(mof_ctx* ctx, const double* a, const double* b, int channels) { return set_signals_common(ctx, a, b, channels, cudaMemcpyHostToDevice); }
int mof_set_signals_device(mof_ctx* ctx, const double* a, const double* b, int channels) { return set_signals_common(ctx, a, b, channels, cudaMemcpyDeviceToDevice); }

int mof_iterate(mof_ctx* ctx, int n) {
    if (!ctx) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    MOF_TRY(require_signals(ctx));
    StreamScope scope(ctx);
    for (int i = 0; i < n; i++) {
        MOF_TRY(update_flow(ctx, ctx->curSmooth, ctx->curVf));
        // IterativeOptimization, OpticalFlow.cpp:1041-1042
        ctx->curSmooth *= ctx->params.sMultiply;
        ctx->curVf = ctx->curVf * ctx->params.vMultiply > ctx->params.vfSThreshold ? ctx->curVf * ctx->params.vMultiply : ctx->curVf;
        ctx->iterationsDone++;
    }
    return MOF_OK;
}

int mof_num_edges(mof_ctx* ctx) { return ctx && ctx->haveMesh ? ctx->E : -1; }
long long mof_num_coeffs(mof_ctx* ctx) { return ctx && ctx->haveMesh ? vf_unknowns(ctx) : -1; }

int mof_get_flow(mof_ctx* ctx, double* tField) {
    if (!ctx || !tField) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    StreamScope scope(ctx);
    const double* field = ctx->tfield.p;
    if (ctx->reordered) {  // a triangle keeps its corner order, hence its chart: only the rows move
        MOF_CUDA(ctx->dtmp2.reserve(2ull * ctx->T));
        MOF_TRY(reorder_scatter(ctx, 1, ctx->tfield.p, 2, ctx->dtmp2.p));
        field = ctx->dtmp2.p;
    }
    MOF_CUDA(cudaMemcpyAsync(tField, field, sizeof(double) * 2 * ctx->T, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}
int mof_get_coeffs(mof_ctx* ctx, double* coeffs) {
    if (!ctx || !coeffs) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    MOF_CUDA(cudaMemcpyAsync(coeffs, ctx->coeffs.p, sizeof(double) * vf_unknowns(ctx), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

static int advect_common(mof_ctx* ctx, double alpha, double* outA, double* outB, cudaMemcpyKind kind) {
    if (!ctx || !outA || !outB) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    MOF_TRY(require_signals(ctx));
    StreamScope scope(ctx);
    const int V = ctx->V;
    // InputGeometryData::flow, OpticalFlow.cpp:482-489: raw colours, lengths -alpha and 1-alpha
    MOF_TRY(advect_vertices(ctx, ctx->raw6.p, -alpha, 1. - alpha, ctx->resampled6.p));
    MOF_CUDA(ctx->dtmp0.reserve(6ull * V));
    double* sa = ctx->dtmp0.p;
    double* sb = ctx->dtmp0.p + 3ull * V;
    MOF_LAUNCH(k_deinterleave, blocks_for(3ll * V, 256), 256, 0, ctx->resampled6.p, V, sa, sb);
    if (ctx->reordered) {  // back to the caller's numbering
        MOF_CUDA(ctx->dtmp2.reserve(6ull * V));
        MOF_TRY(reorder_scatter(ctx, 0, sa, 3, ctx->dtmp2.p));
        MOF_TRY(reorder_scatter(ctx, 0, sb, 3, ctx->dtmp2.p + 3ull * V));
        sa = ctx->dtmp2.p, sb = ctx->dtmp2.p + 3ull * V;
    }
    MOF_CUDA(cudaMemcpyAsync(outA, sa, sizeof(double) * 3 * V, kind, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(outB, sb, sizeof(double) * 3 * V, kind, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}
int mof_advect_vertices(mof_ctx* ctx, double alpha, double* outA, double* outB) { return advect_common(ctx, alpha, outA, outB, cudaMemcpyDeviceToHost); }
int mof_advect_vertices_device(mof_ctx* ctx, double alpha, double* outA, double* outB) { return advect_common(ctx, alpha, outA, outB, cudaMemcpyDeviceToDevice); }

int mof_set_texture_map(mof_ctx* ctx, int W, int H, const int* srcT, const double* srcP, const double* triUV, const unsigned char* texA, const unsigned char* texB) {
    if (!ctx) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    if (W < 2 || H < 2 || !srcT || !srcP || !triUV || !texA || !texB) return fail(ctx, MOF_E_INVALID, "mof_set_texture_map: bad arguments");
    MOF_TRY(require_callers_numbering(ctx, "mof_set_texture_map"));
    StreamScope scope(ctx);
    size_t n = (size_t)W * H;
    for (size_t i = 0; i < n; i++)
        if (srcT[i] < -1 || srcT[i] >= ctx->T) return fail(ctx, MOF_E_INVALID, "mof_set_texture_map: texel refers to a triangle outside the mesh");
    ctx->texW = W, ctx->texH = H;
    MOF_CUDA(ctx->srcT.alloc(n));
    MOF_CUDA(ctx->srcP.alloc(2 * n));
    MOF_CUDA(ctx->triUV.alloc(6ull * ctx->T));
    MOF_CUDA(ctx->tex[0].alloc(3 * n));
    MOF_CUDA(ctx->tex[1].alloc(3 * n));
    MOF_CUDA(cudaMemcpyAsync(ctx->srcT.p, srcT, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->srcP.p, srcP, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->triUV.p, triUV, sizeof(double) * 6 * ctx->T, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->tex[0].p, texA, 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->tex[1].p, texB, 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->haveTexture = true;
    return MOF_OK;
}

int mof_advect_texels(mof_ctx* ctx, double alpha, int bilinear, double* outA, double* outB) {
    if (!ctx || !outA || !outB) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    if (!ctx->haveTexture) return fail(ctx, MOF_E_INVALID, "call mof_set_texture_map first");
    StreamScope scope(ctx);
    MOF_TRY(advect_texels(ctx, alpha, bilinear));
    size_t n = (size_t)ctx->texW * ctx->texH;
    MOF_CUDA(cudaMemcpyAsync(outA, ctx->texOut.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(outB, ctx->texOut.p + 3 * n, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

int mof_advect_texels_frames(mof_ctx* ctx, int frames, int bilinear, double* outA, double* outB) {
    if (!ctx || !outA || !outB) return MOF_E_INVALID;
    if (frames < 2 || frames > 4096) return fail(ctx, MOF_E_INVALID, "mof_advect_texels_frames: 2 <= frames <= 4096");
    MOF_TRY(require_mesh(ctx));
    if (!ctx->haveTexture) return fail(ctx, MOF_E_INVALID, "call mof_set_texture_map first");
    StreamScope scope(ctx);
    MOF_TRY(advect_texels_frames(ctx, frames, bilinear));
    size_t n = (size_t)ctx->texW * ctx->texH * frames;
    MOF_CUDA(cudaMemcpyAsync(outA, ctx->texOut.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(outB, ctx->texOut.p + 3 * n, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

// ---------------------------------------------------- texture-configuration preparation on the device

int mof_subdivide(mof_ctx* ctx, const float* xyz, int V, const int* tri, const double* triUV, int T, double edgeLength, int* outV, int* outT) {
    if (!ctx) return MOF_E_INVALID;
    if (!xyz || !tri || !triUV || V < 1 || T < 1 || !outV || !outT) return fail(ctx, MOF_E_INVALID, "mof_subdivide: bad arguments");
    for (size_t i = 0; i < 3 * (size_t)T; i++)
        if (tri[i] < 0 || tri[i] >= V) return fail(ctx, MOF_E_INVALID, "[ERROR] triangle refers to a vertex index outside [0,V)");
    StreamScope scope(ctx);
    MOF_CUDA(ctx->subXyz.alloc(3ull * V));
    MOF_CUDA(ctx->subTri.alloc(3ull * T));
    MOF_CUDA(ctx->subUv.alloc(6ull * T));
    MOF_CUDA(cudaMemcpyAsync(ctx->subXyz.p, xyz, sizeof(float) * 3 * V, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->subTri.p, tri, sizeof(int) * 3 * T, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->subUv.p, triUV, sizeof(double) * 6 * T, cudaMemcpyHostToDevice, ctx->stream));
    ctx->subV = V, ctx->subT = T;
    if (edgeLength > 0) MOF_TRY(subdivide_mesh(ctx, edgeLength, nullptr));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    *outV = ctx->subV, *outT = ctx->subT;
    return MOF_OK;
}

int mof_get_subdivision(mof_ctx* ctx, float* xyz, int* tri, double* triUV) {
    if (!ctx) return MOF_E_INVALID;
    if (!ctx->subV || !ctx->subXyz.p) return fail(ctx, MOF_E_INVALID, "call mof_subdivide first");
    if (!xyz || !tri || !triUV) return fail(ctx, MOF_E_INVALID, "mof_get_subdivision: bad arguments");
    StreamScope scope(ctx);
    MOF_CUDA(cudaMemcpyAsync(xyz, ctx->subXyz.p, sizeof(float) * 3 * ctx->subV, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(tri, ctx->subTri.p, sizeof(int) * 3 * ctx->subT, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(triUV, ctx->subUv.p, sizeof(double) * 6 * ctx->subT, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

int mof_build_texture_map(mof_ctx* ctx, int W, int H, int padRadius, const double* triUV, const unsigned char* texA, const unsigned char* texB, int* misses) {
    if (!ctx) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    if (W < 2 || H < 2 || (long long)W * H > (1ll << 30) || padRadius < 0 || !triUV || !texA || !texB) return fail(ctx, MOF_E_INVALID, "mof_build_texture_map: bad arguments");
    MOF_TRY(require_callers_numbering(ctx, "mof_build_texture_map"));
    StreamScope scope(ctx);
    size_t n = (size_t)W * H;
    ctx->haveTexture = false;
    ctx->texW = W, ctx->texH = H;
    MOF_CUDA(ctx->triUV.alloc(6ull * ctx->T));
    MOF_CUDA(ctx->tex[0].alloc(3 * n));
    MOF_CUDA(ctx->tex[1].alloc(3 * n));
    MOF_CUDA(cudaMemcpyAsync(ctx->triUV.p, triUV, sizeof(double) * 6 * ctx->T, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->tex[0].p, texA, 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(ctx->tex[1].p, texB, 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    int missed = 0;
    MOF_TRY(build_texture_map(ctx, W, H, padRadius, &missed));
    if (misses) *misses = missed;
    if (missed) {  // the reference exits at the first one (FEM.inl:889)
        char msg[160];
        snprintf(msg, sizeof(msg), "[ERROR] FEM::Mesh::exp:\n        Ray does not intersect triangle (%d texels)", missed);
        return fail(ctx, MOF_E_MESH, msg);
    }
    ctx->haveTexture = true;
    return MOF_OK;
}

int mof_get_texture_map(mof_ctx* ctx, int* srcT, double* srcP) {
    if (!ctx) return MOF_E_INVALID;
    if (!ctx->haveTexture) return fail(ctx, MOF_E_INVALID, "call mof_set_texture_map or mof_build_texture_map first");
    if (!srcT || !srcP) return fail(ctx, MOF_E_INVALID, "mof_get_texture_map: bad arguments");
    StreamScope scope(ctx);
    size_t n = (size_t)ctx->texW * ctx->texH;
    MOF_CUDA(cudaMemcpyAsync(srcT, ctx->srcT.p, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(srcP, ctx->srcP.p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

int mof_sample_textures_to_vertices(mof_ctx* ctx, int bilinear, double* outA, double* outB) {
    if (!ctx) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    if (!ctx->haveTexture) return fail(ctx, MOF_E_INVALID, "call mof_set_texture_map or mof_build_texture_map first");
    if (!outA || !outB) return fail(ctx, MOF_E_INVALID, "mof_sample_textures_to_vertices: bad arguments");
    StreamScope scope(ctx);
    const int V = ctx->V;
    MOF_CUDA(ctx->dtmp0.reserve(6ull * V));
    MOF_CUDA(ctx->dtmp2.reserve(6ull * V));
    MOF_TRY(sample_textures_to_vertices(ctx, bilinear, ctx->dtmp0.p));
    MOF_LAUNCH(k_deinterleave, blocks_for(3ll * V, 256), 256, 0, ctx->dtmp0.p, V, ctx->dtmp2.p, ctx->dtmp2.p + 3ull * V);
    MOF_CUDA(cudaMemcpyAsync(outA, ctx->dtmp2.p, sizeof(double) * 3 * V, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(outB, ctx->dtmp2.p + 3ull * V, sizeof(double) * 3 * V, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

// ------------------------------------------------------------------------------------ debug taps

static int csr_view(mof_ctx* ctx, int which, int* rows, long long* nnz, const int** rowptr, const int** col, const double** val) {
    MOF_TRY(require_mesh(ctx));
    switch (which) {
        case MOF_CSR_SCALAR_MASS: *rows = ctx->V, *nnz = ctx->nnzS, *rowptr = ctx->sRowptr.p, *col = ctx->sCol.p, *val = ctx->sMass.p; return MOF_OK;
        case MOF_CSR_SCALAR_STIFFNESS: *rows = ctx->V, *nnz = ctx->nnzS, *rowptr = ctx->sRowptr.p, *col = ctx->sCol.p, *val = ctx->sStiff.p; return MOF_OK;
        case MOF_CSR_FLOW_SYSTEM:
            if (vf_active(ctx)) return fail(ctx, MOF_E_UNSUPPORTED, "the Conformal / Connection flow systems are applied matrix-free: there is no assembled matrix to return");
            break;
        default: break;
    }
    switch (which) {
        case MOF_CSR_WHITNEY_SMOOTH: *rows = ctx->E, *nnz = ctx->nnzW, *rowptr = ctx->wRowptr.p, *col = ctx->wCol.p, *val = ctx->wS.p; return MOF_OK;
        case MOF_CSR_FLOW_SYSTEM:
            if (!ctx->haveFlowSystem) return fail(ctx, MOF_E_INVALID, "no flow system yet: call mof_iterate first");
            *rows = ctx->E, *nnz = ctx->nnzW, *rowptr = ctx->wRowptr.p, *col = ctx->wCol.p, *val = ctx->wA.p;
            return MOF_OK;
    }
    return fail(ctx, MOF_E_INVALID, "unknown CSR id");
}

int mof_csr_size(mof_ctx* ctx, int which, int* rows, long long* nnz) {
    if (!ctx || !rows || !nnz) return MOF_E_INVALID;
    const int *rp, *c;
    const double* v;
    return csr_view(ctx, which, rows, nnz, &rp, &c, &v);
}

int mof_get_csr(mof_ctx* ctx, int which, int* rowptr, int* col, double* val) {
    if (!ctx || !rowptr || !col || !val) return MOF_E_INVALID;
    int rows;
    long long nnz;
    const int *rp, *c;
    const double* v;
    MOF_TRY(csr_view(ctx, which, &rows, &nnz, &rp, &c, &v));
    StreamScope scope(ctx);
    DBuf<int> tmpCol;
    DBuf<double> tmpVal;
    if (which == MOF_CSR_WHITNEY_SMOOTH || which == MOF_CSR_FLOW_SYSTEM) {
        // the E x E operators live in the sliced layout: unpack to CSR for the caller
        MOF_CUDA(tmpCol.alloc((size_t)nnz));
        MOF_CUDA(tmpVal.alloc((size_t)nnz));
        int rc = sell_to_csr(ctx, rows, rp, ctx->wSliceBase.p, c, v, tmpCol.p, tmpVal.p);
        if (rc != MOF_OK) { tmpCol.release(), tmpVal.release(); return rc; }
        c = tmpCol.p, v = tmpVal.p;
    }
    cudaError_t e = cudaMemcpyAsync(rowptr, rp, sizeof(int) * (rows + 1), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(col, c, sizeof(int) * nnz, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(val, v, sizeof(double) * nnz, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    tmpCol.release(), tmpVal.release();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "mof_get_csr copy");
    return MOF_OK;
}

static const void* array_view(mof_ctx* ctx, int which, long long* bytes) {
    const long long V = ctx->V, T = ctx->T, E = ctx->E;
    *bytes = 0;
    if (!ctx->haveMesh) return nullptr;
    switch (which) {
        case MOF_ARR_METRIC: *bytes = 8 * 3 * T; return ctx->g.p;
        case MOF_ARR_AREA: *bytes = 8 * T; return ctx->area.p;
        case MOF_ARR_OPPOSITE: *bytes = 4 * 3 * T; return ctx->opp.p;
        case MOF_ARR_XFORM_LINEAR: *bytes = 8 * 12 * T; return ctx->xlin.p;
        case MOF_ARR_XFORM_CONSTANT: *bytes = 8 * 6 * T; return ctx->xcst.p;
        case MOF_ARR_REDUCED_EDGE: *bytes = 4 * 3 * T; return ctx->reduced.p;
        case MOF_ARR_EXPANDED_EDGE: *bytes = 4 * E; return ctx->expanded.p;
        case MOF_ARR_POSITIVE_EDGE: *bytes = 4 * 3 * T; return ctx->positive.p;
        case MOF_ARR_PROLONGATION: *bytes = 8 * 6 * T; return ctx->P.p;
        default: break;
    }
    if (!ctx->haveSignals) return nullptr;
    switch (which) {
        case MOF_ARR_SIGNALS: *bytes = 8 * 6 * V; return ctx->sig6.p;
        case MOF_ARR_SIGNALS_RAW:
            if (!ctx->blend) return nullptr;
            *bytes = 8 * 6 * V;
            return ctx->sigLo6.p;
        default: break;
    }
    if (!ctx->haveFlowSystem) return nullptr;
    switch (which) {
        case MOF_ARR_SMOOTHED: *bytes = 8 * 6 * V; return ctx->smoothed6.p;
        case MOF_ARR_RESAMPLED: *bytes = 8 * 6 * V; return ctx->resampled6.p;
        case MOF_ARR_RESAMPLED_RAW:
            if (!ctx->blend) return nullptr;
            *bytes = 8 * 6 * V;
            return ctx->resampledLo6.p;
        case MOF_ARR_DATA_TERM: *bytes = 8 * 3 * T; return ctx->dataD.p;
        case MOF_ARR_DATA_RHS: *bytes = 8 * 2 * T; return ctx->dataRhs.p;
        case MOF_ARR_FLOW_RHS: *bytes = 8 * vf_unknowns(ctx); return vf_active(ctx) ? vf_rhs(ctx) : ctx->fb.p;
        case MOF_ARR_FLOW_SOLUTION: *bytes = 8 * vf_unknowns(ctx); return vf_active(ctx) ? vf_solution(ctx) : ctx->fx.p;
        default: break;
    }
    return nullptr;
}

long long mof_array_bytes(mof_ctx* ctx, int which) {
    if (!ctx) return 0;
    long long bytes = 0;
    array_view(ctx, which, &bytes);
    return bytes;
}

int mof_get_array(mof_ctx* ctx, int which, void* out) {
    if (!ctx || !out) return MOF_E_INVALID;
    long long bytes = 0;
    const void* src = array_view(ctx, which, &bytes);
    if (!src || !bytes) return fail(ctx, MOF_E_INVALID, "array not available");
    MOF_CUDA(cudaMemcpyAsync(out, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

int mof_pcg_solve_csr(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, const double* b, double* x, double tol, int maxIters, int* iters,
                      double* relres) {
    if (!ctx || n < 1 || !rowptr || !col || !val || !b || !x || !iters || !relres) return MOF_E_INVALID;
    StreamScope scope(ctx);
    long long nnz = rowptr[n];
    DBuf<int> dRow, dCol, sBase, sCol;
    DBuf<double> dVal, dB, dX, dInv, sVal;
    int rc = MOF_OK;
    auto cleanup = [&]() {
        dRow.release(), dCol.release(), dVal.release(), dB.release(), dX.release(), dInv.release();
        sBase.release(), sCol.release(), sVal.release();
    };
    cudaError_t e;
    if ((e = dRow.alloc(n + 1)) != cudaSuccess || (e = dCol.alloc(nnz)) != cudaSuccess || (e = dVal.alloc(nnz)) != cudaSuccess || (e = dB.alloc(n)) != cudaSuccess ||
        (e = dX.alloc(n)) != cudaSuccess || (e = dInv.alloc(n)) != cudaSuccess) {
        cleanup();
        return cuda_fail(ctx, e, "mof_pcg_solve_csr alloc");
    }
    cudaMemcpyAsync(dRow.p, rowptr, sizeof(int) * (n + 1), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dCol.p, col, sizeof(int) * nnz, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dVal.p, val, sizeof(double) * nnz, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dB.p, b, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    rc = extract_inverse_diagonal(ctx, n, dRow.p, dCol.p, dVal.p, dInv.p);
    if (rc == MOF_OK) rc = csr_to_sell(ctx, n, dRow.p, dCol.p, dVal.p, sBase, sCol, sVal);
    if (rc == MOF_OK) rc = pcg_solve_sell(ctx, n, sBase.p, sCol.p, sVal.p, dInv.p, dB.p, dX.p, true, tol, maxIters, iters, relres);
    if (rc == MOF_OK || rc == MOF_E_NOCONVERGE) {
        cudaMemcpyAsync(x, dX.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    cleanup();
    return rc;
}

int mof_time_flow_spmv(mof_ctx* ctx, int reps, float* msPerLaunch) {
    if (!ctx || reps < 1 || !msPerLaunch) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    if (!ctx->haveFlowSystem) return fail(ctx, MOF_E_INVALID, "no flow system yet: call mof_iterate first");
    if (vf_active(ctx)) return fail(ctx, MOF_E_UNSUPPORTED, "mof_time_flow_spmv times the Whitney flow matrix");
    StreamScope scope(ctx);
    MOF_CUDA(ctx->pcg.q.reserve(ctx->E));
    return time_spmv_sell(ctx, ctx->E, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, ctx->fx.p, ctx->pcg.q.p, reps, msPerLaunch);
}

int mof_time_kernel(mof_ctx* ctx, int which, int reps, double* usPerLaunch, double* algorithmicBytes) {
    if (!ctx || reps < 1 || !usPerLaunch || !algorithmicBytes || which < 0 || which >= MOF_K_COUNT) return MOF_E_INVALID;
    MOF_TRY(require_mesh(ctx));
    if (!ctx->haveFlowSystem) return fail(ctx, MOF_E_INVALID, "no flow system yet: call mof_iterate first");
    if (vf_active(ctx)) return fail(ctx, MOF_E_UNSUPPORTED, "mof_time_kernel times the kernels of the Whitney basis");
    if (dist_active(ctx)) return fail(ctx, MOF_E_UNSUPPORTED, "mof_time_kernel: not on a partitioned mesh");
    StreamScope scope(ctx);
    smooth_ahead_drain(ctx);
    float ms = 0;
    double bytes = 0;
    if (which == MOF_K_WALK) MOF_TRY(time_walk_kernel(ctx, reps, &ms));
    else MOF_TRY(mg_time_kernel(ctx, which, reps, &ms, &bytes));
    *usPerLaunch = ms * 1e3, *algorithmicBytes = bytes;
    return MOF_OK;
}

}  // extern "C"
