// TEST INFRASTRUCTURE (CPU tier): meshopticalflow_b200/csrc/setup_kernels.cu — operator assembly on the "device": metric,
// half-edge adjacency through the open-addressing table, edge transforms, scalar mass / stiffness CSR, Whitney
// numbering, prolongation and smoothness operator in the sliced layout, the scans and reductions they use — kernels AND
// host driver (build_mesh_operators), the very source the GPU build compiles, built for the host through
// emul_cuda_runtime.h (fibers; counted barriers, warp shuffles and atomics emulated).
#include "emul_cuda_runtime.h"

#include "../../meshopticalflow_b200/csrc/setup_kernels.cu"

namespace {
mof_ctx* g_ctx = nullptr;
template <class T>
void adopt(mof::DBuf<T>& b, const T* host, size_t n) {
    b.alloc(n);
    memcpy(b.p, host, n * sizeof(T));
}
}  // namespace

extern "C" {

// mof_set_mesh's device part on host arrays. sizes: [E, nnzS, nnzW, wSlices, wPadded]. Returns the library's status code;
// the message of a failure is left in `message`.
int emul_mesh_build(int V, int T, const double* xyz, const int* tri, long long* sizes, char* message, int messageBytes) {
    delete g_ctx;
    g_ctx = new mof_ctx();
    mof_ctx* ctx = g_ctx;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->V = V, ctx->T = T;
    adopt(ctx->pos, xyz, 3 * (size_t)V), adopt(ctx->tri, tri, 3 * (size_t)T);
    int rc = mof::build_mesh_operators(ctx);
    if (message && messageBytes > 0) snprintf(message, messageBytes, "%s", ctx->err.c_str());
    if (rc != MOF_OK) return rc;
    sizes[0] = ctx->E, sizes[1] = ctx->nnzS, sizes[2] = ctx->nnzW, sizes[3] = ctx->wSlices, sizes[4] = ctx->wPadded;
    return MOF_OK;
}

// which: the names below; copies the whole buffer.
int emul_mesh_get(const char* which, void* out) {
    mof_ctx* c = g_ctx;
    if (!c) return MOF_E_INVALID;
    struct Item { const char* name; const void* p; size_t bytes; };
    const Item items[] = {
        {"g", c->g.p, c->g.bytes()}, {"area", c->area.p, c->area.bytes()}, {"opp", c->opp.p, c->opp.bytes()}, {"xlin", c->xlin.p, c->xlin.bytes()},
        {"xcst", c->xcst.p, c->xcst.bytes()}, {"sRowptr", c->sRowptr.p, c->sRowptr.bytes()}, {"sCol", c->sCol.p, c->sCol.bytes()}, {"sHe", c->sHe.p, c->sHe.bytes()},
        {"sMass", c->sMass.p, c->sMass.bytes()}, {"sStiff", c->sStiff.p, c->sStiff.bytes()}, {"m0", c->m0.p, c->m0.bytes()}, {"reduced", c->reduced.p, c->reduced.bytes()},
        {"expanded", c->expanded.p, c->expanded.bytes()}, {"positive", c->positive.p, c->positive.bytes()}, {"P", c->P.p, c->P.bytes()},
        {"wRowptr", c->wRowptr.p, c->wRowptr.bytes()}, {"wSliceBase", c->wSliceBase.p, c->wSliceBase.bytes()}, {"wCol", c->wCol.p, c->wCol.bytes()},
        {"wS", c->wS.p, c->wS.bytes()},
    };
    for (const Item& it : items)
        if (!strcmp(it.name, which)) {
            memcpy(out, it.p, it.bytes);
            return MOF_OK;
        }
    return MOF_E_INVALID;
}

}  // extern "C"
