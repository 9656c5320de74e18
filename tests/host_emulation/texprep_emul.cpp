// TEST INFRASTRUCTURE (CPU tier): meshopticalflow_b200/csrc/texprep_kernels.cu — the texture configuration's one-time
// preparation on the "device" (edge-length subdivision, texel -> (triangle, point) map, wedge-averaged vertex colours) —
// kernels AND host drivers, the very source the GPU build compiles, built for the host through emul_cuda_runtime.h,
// on top of setup_kernels.cu (the scans, and the edge transforms / vertex rows the map and the sampling read).
#include "emul_cuda_runtime.h"

#include "../../meshopticalflow_b200/csrc/setup_kernels.cu"
#include "../../meshopticalflow_b200/csrc/texprep_kernels.cu"

namespace {
mof_ctx* g_ctx = nullptr;
template <class T>
void adopt(mof::DBuf<T>& b, const T* host, size_t n) {
    b.alloc(n);
    memcpy(b.p, host, n * sizeof(T));
}
mof_ctx* fresh() {
    delete g_ctx;
    g_ctx = new mof_ctx();
    memset(&g_ctx->stats, 0, sizeof(g_ctx->stats));
    return g_ctx;
}
}  // namespace

extern "C" {

// mof_subdivide's device part. sizes: [V, T] after subdivision.
int emul_subdivide(int V, int T, const float* xyz, const int* tri, const double* uv, double edgeLength, int* sizes) {
    mof_ctx* ctx = fresh();
    adopt(ctx->subXyz, xyz, 3 * (size_t)V), adopt(ctx->subTri, tri, 3 * (size_t)T), adopt(ctx->subUv, uv, 6 * (size_t)T);
    ctx->subV = V, ctx->subT = T;
    int rc = edgeLength > 0 ? mof::subdivide_mesh(ctx, edgeLength, nullptr) : MOF_OK;
    sizes[0] = ctx->subV, sizes[1] = ctx->subT;
    return rc;
}
int emul_get_subdivision(float* xyz, int* tri, double* uv) {
    mof_ctx* c = g_ctx;
    memcpy(xyz, c->subXyz.p, sizeof(float) * 3 * c->subV), memcpy(tri, c->subTri.p, sizeof(int) * 3 * c->subT), memcpy(uv, c->subUv.p, sizeof(double) * 6 * c->subT);
    return MOF_OK;
}

// mof_set_mesh's operator assembly, then mof_build_texture_map's and mof_sample_textures_to_vertices' device parts.
int emul_texture_prepare(int V, int T, const double* xyz, const int* tri, const double* triUV, int W, int H, int pad, const unsigned char* texA,
                         const unsigned char* texB, int bilinear, int* srcT, double* srcP, double* colors6, int* misses, long long* launches) {
    mof_ctx* ctx = fresh();
    ctx->V = V, ctx->T = T;
    adopt(ctx->pos, xyz, 3 * (size_t)V), adopt(ctx->tri, tri, 3 * (size_t)T);
    int rc = mof::build_mesh_operators(ctx);
    if (rc != MOF_OK) return rc;
    size_t n = (size_t)W * H;
    adopt(ctx->triUV, triUV, 6 * (size_t)T), adopt(ctx->tex[0], texA, 3 * n), adopt(ctx->tex[1], texB, 3 * n);
    ctx->texW = W, ctx->texH = H;
    long long before = ctx->stats.kernelLaunches;
    rc = mof::build_texture_map(ctx, W, H, pad, misses);
    if (rc != MOF_OK) return rc;
    memcpy(srcT, ctx->srcT.p, sizeof(int) * n), memcpy(srcP, ctx->srcP.p, sizeof(double) * 2 * n);
    mof::DBuf<double> out;
    out.alloc(6 * (size_t)V);
    rc = mof::sample_textures_to_vertices(ctx, bilinear, out.p);
    if (rc != MOF_OK) return rc;
    memcpy(colors6, out.p, sizeof(double) * 6 * V);
    out.release();
    *launches = ctx->stats.kernelLaunches - before;
    return MOF_OK;
}

}  // extern "C"
