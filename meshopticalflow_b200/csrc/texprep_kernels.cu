// Texture-configuration preparation on the GPU: edge-length subdivision, wedge-averaged
// vertex colours and the texel -> (triangle, barycentric point) map of include/Src/MeshFlow.inl:158-232, 252-266,
// 281-467 — the one-time work the reference (and this build's host/texture_prep.cpp) does serially on the CPU.
//
// The reference's loops are sequential and order dependent (midpoint vertices are numbered in the order their edges
// are first met, a texel keeps the FIRST triangle that covers it, a vertex colour is summed wedge by wedge in
// triangle order). Every kernel below reproduces that order by a gather with a fixed rule instead of by running
// serially:
//   * midpoint of an edge: its rank among the "first" half-edges (the smaller 3t+j of the edge's two), by a scan;
//   * owner of a texel: the LARGEST triangle index whose point passes the rule of MeshFlow.inl:334, else the SMALLEST
//     covering triangle (two atomics per covered texel, then a second pass in which only the owner writes);
//   * vertex colour: the incident wedges visited in ascending triangle order.
// All floating-point arithmetic that decides a branch or lands in an output goes through dmul/dadd/dsub (__dmul_rn
// ...), which the compiler never contracts into FMAs: the results are bit-identical to the host code's IEEE double
// evaluation, so the integer outputs (triangle per texel, connectivity, numbering) match exactly, not "almost".
#include <climits>

#include "mof_internal.cuh"

namespace mof {

#ifdef MOF_HOST_EMULATION
static inline double dmul(double a, double b) { return a * b; }
static inline double dadd(double a, double b) { return a + b; }
static inline double dsub(double a, double b) { return a - b; }
#else
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
#endif

struct V2 {
    double x, y;
};
__device__ __forceinline__ V2 v2(double x, double y) { V2 r; r.x = x, r.y = y; return r; }
__device__ __forceinline__ V2 vsub(V2 a, V2 b) { return v2(dsub(a.x, b.x), dsub(a.y, b.y)); }

// BarycentricCoordinate, MeshFlow.inl:268-278: p = c0 + s (c1 - c0) + t (c2 - c0).
__device__ __forceinline__ V2 barycentric(const V2* c, V2 p) {
    V2 a = vsub(c[1], c[0]), b = vsub(c[2], c[0]), r = vsub(p, c[0]);
    double inv = 1. / dsub(dmul(a.x, b.y), dmul(b.x, a.y));
    return v2(dmul(dsub(dmul(b.y, r.x), dmul(b.x, r.y)), inv), dmul(dadd(dmul(-a.y, r.x), dmul(a.x, r.y)), inv));
}

__device__ __forceinline__ int clampi(int v, int n) { return max(0, min(n - 1, v)); }

// ------------------------------------------------------------------------------- rasterisation

constexpr int RASTER_LANES = 8;  // lanes that share one triangle (they split a scan line between them)

// RasterizeTriangle, MeshFlow.inl:281-337: corners ordered by v, scan lines between the two active sides, the point of a
// texel interpolated between the points of the scan line's two ends. PASS 0 records for every covered texel the
// smallest covering triangle and the largest one that passes the rule of :334; PASS 1 lets the owner write.
template <int PASS>
__global__ void k_raster(const double* __restrict__ triUV, int T, int W, int H, int* __restrict__ first, int* __restrict__ lastRule, int* __restrict__ srcT,
                         double* __restrict__ srcP) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int t = (int)(gid / RASTER_LANES), lane = (int)(gid % RASTER_LANES);
    if (t >= T) return;
    const double* uv = triUV + 6 * (size_t)t;
    V2 c[3];
    for (int j = 0; j < 3; j++) c[j] = v2(dmul(uv[2 * j], (double)(W - 1)), dmul(uv[2 * j + 1], (double)(H - 1)));
    int o0, o1, o2;
    double y0 = uv[1], y1 = uv[3], y2 = uv[5];
    if (y0 <= y1 && y0 <= y2) o0 = 0, o1 = y1 <= y2 ? 1 : 2, o2 = y1 <= y2 ? 2 : 1;
    else if (y1 <= y0 && y1 <= y2) o0 = 1, o1 = y0 <= y2 ? 0 : 2, o2 = y0 <= y2 ? 2 : 0;
    else o0 = 2, o1 = y0 <= y1 ? 0 : 1, o2 = y0 <= y1 ? 1 : 0;
    V2 lo = c[o0], mid = c[o1], hi = c[o2];
    int yBegin = clampi((int)ceil(lo.y), H), yEnd = clampi((int)floor(hi.y), H);
    V2 apex = lo, sideA = vsub(mid, lo), sideB = vsub(hi, lo);
    for (int y = yBegin; y <= yEnd; y++) {
        if ((double)y >= mid.y) apex = hi, sideA = vsub(mid, hi), sideB = vsub(lo, hi);
        if (sideA.y == 0 || sideB.y == 0) continue;
        double dy = dsub((double)y, apex.y);
        double xa = dadd(apex.x, dmul(dy, sideA.x) / sideA.y), xb = dadd(apex.x, dmul(dy, sideB.x) / sideB.y);
        int xBegin = clampi((int)ceil(fmin(xa, xb)), W), xEnd = clampi((int)floor(fmax(xa, xb)), W);
        V2 bBegin = barycentric(c, v2((double)xBegin, (double)y)), bEnd = barycentric(c, v2((double)xEnd, (double)y));
        for (int x = xBegin + lane; x <= xEnd; x += RASTER_LANES) {
            double s = xBegin == xEnd ? 0. : (double)(x - xBegin) / (double)(xEnd - xBegin);
            double r = dsub(1., s);
            V2 b = v2(dadd(dmul(bBegin.x, r), dmul(bEnd.x, s)), dadd(dmul(bBegin.y, r), dmul(bEnd.y, s)));
            size_t i = (size_t)y * W + x;
            if (PASS == 0) {
                atomicMin(&first[i], t);
                if (b.x >= 0 && b.y >= 1 && dadd(b.x, b.y) <= 1) atomicMax(&lastRule[i], t);
            } else {
                int owner = lastRule[i] >= 0 ? lastRule[i] : first[i];
                if (owner == t) srcT[i] = t, srcP[2 * i] = b.x, srcP[2 * i + 1] = b.y;
            }
        }
    }
}

__global__ void k_texmap_clear(int n, int* __restrict__ first, int* __restrict__ lastRule, int* __restrict__ srcT, double* __restrict__ srcP) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    first[i] = INT_MAX, lastRule[i] = -1, srcT[i] = -1, srcP[2 * i] = 0., srcP[2 * i + 1] = 0.;
}

// One padding ring, MeshFlow.inl:426-455, first half: an empty texel takes the triangle of a covered 4-neighbour — the
// last one found in the reference's order x-1, x+1, y-1, y+1. The ring reads the map as it was before the ring.
__global__ void k_pad_find(const int* __restrict__ srcT, int W, int H, int* __restrict__ grow) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    int y = i / W, x = i - y * W, found = -1;
    if (srcT[i] == -1) {
        if (x - 1 >= 0 && srcT[i - 1] != -1) found = srcT[i - 1];
        if (x + 1 < W && srcT[i + 1] != -1) found = srcT[i + 1];
        if (y - 1 >= 0 && srcT[i - W] != -1) found = srcT[i - W];
        if (y + 1 < H && srcT[i + W] != -1) found = srcT[i + W];
    }
    grow[i] = found;
}

// ... second half: the texel's point in the chart of that triangle (outside it, by construction).
__global__ void k_pad_apply(const int* __restrict__ grow, const double* __restrict__ triUV, int W, int H, int* __restrict__ srcT, double* __restrict__ srcP) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    int t = grow[i];
    if (t == -1) return;
    int y = i / W, x = i - y * W;
    const double* uv = triUV + 6 * (size_t)t;
    V2 c[3];
    for (int j = 0; j < 3; j++) c[j] = v2(uv[2 * j], uv[2 * j + 1]);
    V2 b = barycentric(c, v2((double)x / (double)(W - 1), (double)y / (double)(H - 1)));
    srcT[i] = t, srcP[2 * i] = b.x, srcP[2 * i + 1] = b.y;
}

// RiemannianMesh::exp, FEM.inl:835-899: the straight line from p with velocity v, across edges, until the velocity is
// used up. Returns false when the ray misses its triangle (the reference exits there).
__device__ bool exp_map(const int* __restrict__ opp, const double* __restrict__ xlin, const double* __restrict__ xcst, int& t, V2& p, V2 v) {
    if (!dadd(dmul(v.x, v.x), dmul(v.y, v.y))) return true;
    int cameFrom = -1;
    auto cross = [&](int side) {
        size_t h = 3 * (size_t)t + side;
        int o = opp[h];
        const double* L = xlin + 4 * h;
        const double* k = xcst + 2 * h;
        p = v2(dadd(dadd(dmul(L[0], p.x), dmul(L[1], p.y)), k[0]), dadd(dadd(dmul(L[2], p.x), dmul(L[3], p.y)), k[1]));
        v = v2(dadd(dmul(L[0], v.x), dmul(L[1], v.y)), dadd(dmul(L[2], v.x), dmul(L[3], v.y)));
        t = o / 3, cameFrom = o % 3;
    };
    if (p.x <= 0 && v.x < 0) cross(1);
    else if (p.y <= 0 && v.y < 0) cross(2);
    else if (dadd(p.x, p.y) >= 1 && dadd(v.x, v.y) > 0) cross(0);
    for (int count = 0; count < 10000; count++) {
        double best = 0;
        int side = -1;
        double s2 = -p.y / v.y, s1 = -p.x / v.x, s0 = dsub(dsub(1., p.x), p.y) / dadd(v.y, v.x);
        if (cameFrom != 2 && s2 > 0) { double q = dadd(p.x, dmul(v.x, s2)); if (q >= 0 && q <= 1 && s2 > best) side = 2, best = s2; }
        if (cameFrom != 1 && s1 > 0) { double q = dadd(p.y, dmul(v.y, s1)); if (q >= 0 && q <= 1 && s1 > best) side = 1, best = s1; }
        if (cameFrom != 0 && s0 > 0) { double q = dadd(p.x, dmul(v.x, s0)); if (q >= 0 && q <= 1 && s0 > best) side = 0, best = s0; }
        if (side == -1) return false;
        if (best > 1) {
            p = v2(dadd(p.x, v.x), dadd(p.y, v.y));
            return true;
        }
        p = v2(dadd(p.x, dmul(v.x, best)), dadd(p.y, dmul(v.y, best)));
        v = v2(dsub(v.x, dmul(v.x, best)), dsub(v.y, dmul(v.y, best)));
        cross(side);
    }
    return true;  // "[WARNING] Failed to converge exp" in the reference
}

// RemapSamplePoint, MeshFlow.inl:340-350: a point outside its triangle is reached from the centroid along the surface.
__global__ void k_remap(const int* __restrict__ opp, const double* __restrict__ xlin, const double* __restrict__ xcst, int n, int* __restrict__ srcT,
                        double* __restrict__ srcP, int* __restrict__ misses) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = srcT[i];
    if (t == -1) return;
    V2 p = v2(srcP[2 * i], srcP[2 * i + 1]);
    if (p.x >= 0 && p.y >= 0 && dadd(p.x, p.y) <= 1) return;
    V2 start = v2(1. / 3, 1. / 3), q = start;
    if (!exp_map(opp, xlin, xcst, t, q, vsub(p, start))) atomicAdd(misses, 1);
    srcT[i] = t, srcP[2 * i] = q.x, srcP[2 * i + 1] = q.y;
}

// GetTextureSource, MeshFlow.inl:411-467, from ctx->triUV and the edge transforms of the mesh in place; fills ctx->srcT /
// ctx->srcP (the arrays mof_set_texture_map would have uploaded).
int build_texture_map(mof_ctx* ctx, int W, int H, int padRadius, int* missesOut) {
    const int n = W * H, B = 256, T = ctx->T;
    MOF_CUDA(ctx->srcT.alloc(n));
    MOF_CUDA(ctx->srcP.alloc(2ull * n));
    MOF_CUDA(ctx->itmp0.reserve(n));
    MOF_CUDA(ctx->itmp1.reserve(n));
    MOF_CUDA(ctx->flags.alloc(16));
    MOF_CUDA(cudaMemsetAsync(ctx->flags.p, 0, ctx->flags.bytes(), ctx->stream));
    int *first = ctx->itmp0.p, *lastRule = ctx->itmp1.p;
    MOF_LAUNCH(k_texmap_clear, blocks_for(n, B), B, 0, n, first, lastRule, ctx->srcT.p, ctx->srcP.p);
    const int rasterBlocks = blocks_for((long long)T * RASTER_LANES, B);
    MOF_LAUNCH(k_raster<0>, rasterBlocks, B, 0, ctx->triUV.p, T, W, H, first, lastRule, ctx->srcT.p, ctx->srcP.p);
    MOF_LAUNCH(k_raster<1>, rasterBlocks, B, 0, ctx->triUV.p, T, W, H, first, lastRule, ctx->srcT.p, ctx->srcP.p);
    for (int ring = 0; ring < padRadius; ring++) {
        MOF_LAUNCH(k_pad_find, blocks_for(n, B), B, 0, ctx->srcT.p, W, H, first);
        MOF_LAUNCH(k_pad_apply, blocks_for(n, B), B, 0, first, ctx->triUV.p, W, H, ctx->srcT.p, ctx->srcP.p);
    }
    MOF_LAUNCH(k_remap, blocks_for(n, B), B, 0, ctx->opp.p, ctx->xlin.p, ctx->xcst.p, n, ctx->srcT.p, ctx->srcP.p, ctx->flags.p);
    int misses = 0;
    MOF_CUDA(read_back(ctx, &misses, ctx->flags.p));
    if (missesOut) *missesOut = misses;
    return MOF_OK;
}

// ------------------------------------------------------------------------- texture -> vertices

// Sample, MeshFlow.inl:66-84 (RGB8 texture, rows top to bottom, v up). The four products are added left to right.
__device__ __forceinline__ void sample_rgb8(const unsigned char* __restrict__ tex, int W, int H, double u, double v, int bilinear, double* rgb) {
    v = dsub(1., v);
    u = dmul(fmin(1., fmax(0., u)), (double)(W - 1));
    v = dmul(fmin(1., fmax(0., v)), (double)(H - 1));
    int x0 = (int)floor(u), y0 = (int)floor(v);
    if (!bilinear) {
        for (int c = 0; c < 3; c++) rgb[c] = (double)tex[3 * ((size_t)W * y0 + x0) + c];
        return;
    }
    double dx = dsub(u, (double)x0), dy = dsub(v, (double)y0);
    int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
    double w00 = dmul(dsub(1., dx), dsub(1., dy)), w10 = dmul(dx, dsub(1., dy)), w11 = dmul(dx, dy), w01 = dmul(dsub(1., dx), dy);
    for (int c = 0; c < 3; c++) {
        double p00 = (double)tex[3 * ((size_t)W * y0 + x0) + c], p10 = (double)tex[3 * ((size_t)W * y0 + x1) + c];
        double p11 = (double)tex[3 * ((size_t)W * y1 + x1) + c], p01 = (double)tex[3 * ((size_t)W * y1 + x0) + c];
        rgb[c] = dadd(dadd(dadd(dmul(p00, w00), dmul(p10, w10)), dmul(p11, w11)), dmul(p01, w01));
    }
}

// SampleTextureToVertices, MeshFlow.inl:252-266: a vertex is the mean of its wedges' samples. The wedges of vertex a are
// the triangles behind its outgoing half-edges (row a of the scalar pattern: sHe = half-edge a -> b, which leaves corner
// (j+1)%3 of triangle h/3), visited in ascending triangle order like the reference's triangle loop.
__global__ void k_sample_vertices(const int* __restrict__ sRowptr, const int* __restrict__ sHe, const double* __restrict__ triUV, int V,
                                  const unsigned char* __restrict__ texA, const unsigned char* __restrict__ texB, int W, int H, int bilinear,
                                  double* __restrict__ out6) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= V) return;
    double sum[6] = {0, 0, 0, 0, 0, 0};
    int wedges = 0, lastT = -1;
    const int r0 = sRowptr[a], r1 = sRowptr[a + 1];
    while (true) {
        int bestT = INT_MAX, bestCorner = 0;
        for (int k = r0; k < r1; k++) {
            int h = sHe[k];
            if (h < 0) continue;
            int t = h / 3;
            if (t > lastT && t < bestT) bestT = t, bestCorner = (h - 3 * t + 1) % 3;
        }
        if (bestT == INT_MAX) break;
        lastT = bestT;
        const double* uv = triUV + 6 * (size_t)bestT + 2 * bestCorner;
        double rgb[3];
        sample_rgb8(texA, W, H, uv[0], uv[1], bilinear, rgb);
        for (int c = 0; c < 3; c++) sum[c] = dadd(sum[c], rgb[c]);
        sample_rgb8(texB, W, H, uv[0], uv[1], bilinear, rgb);
        for (int c = 0; c < 3; c++) sum[3 + c] = dadd(sum[3 + c], rgb[c]);
        wedges++;
    }
    for (int c = 0; c < 6; c++) out6[6 * (size_t)a + c] = sum[c] / (double)wedges;
}

int sample_textures_to_vertices(mof_ctx* ctx, int bilinear, double* d_out6) {
    const int B = 128;
    MOF_LAUNCH(k_sample_vertices, blocks_for(ctx->V, B), B, 0, ctx->sRowptr.p, ctx->sHe.p, ctx->triUV.p, ctx->V, ctx->tex[0].p, ctx->tex[1].p, ctx->texW, ctx->texH,
               bilinear, d_out6);
    return MOF_OK;
}

// --------------------------------------------------------------------------------- subdivision

__device__ __forceinline__ unsigned long long edge_key(int a, int b) {
    return a > b ? ((unsigned long long)(unsigned)a << 32) | (unsigned)b : ((unsigned long long)(unsigned)b << 32) | (unsigned)a;
}
__device__ __forceinline__ unsigned edge_slot(unsigned long long k, unsigned mask) { return (unsigned)((k * 0x9E3779B97F4A7C15ull) >> 32) & mask; }
constexpr unsigned long long EDGE_EMPTY = ~0ull;

// _Subdivide, MeshFlow.inl:158-220, step 1. Side j of a triangle runs corner j -> corner (j+1)%3. A side longer than the
// threshold enters the table of undirected edges; the table keeps the smallest side index 3t+j of each edge — the side at
// which the reference's scan meets the edge first and creates its midpoint vertex.
__global__ void k_sub_mark(const float* __restrict__ xyz, const int* __restrict__ tri, int nSides, double threshold2, unsigned long long* keys, int* firstSide,
                           unsigned mask, int* __restrict__ slotOf) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nSides) return;
    int t = h / 3, j = h - 3 * t;
    int a = tri[3 * t + j], b = tri[3 * t + (j + 1) % 3];
    float dx = xyz[3 * (size_t)a] - xyz[3 * (size_t)b], dy = xyz[3 * (size_t)a + 1] - xyz[3 * (size_t)b + 1], dz = xyz[3 * (size_t)a + 2] - xyz[3 * (size_t)b + 2];
    double len2 = dadd(dadd(dmul((double)dx, (double)dx), dmul((double)dy, (double)dy)), dmul((double)dz, (double)dz));
    if (!(len2 > threshold2)) {
        slotOf[h] = -1;
        return;
    }
    unsigned long long key = edge_key(a, b);
    unsigned slot = edge_slot(key, mask);
    while (true) {
        unsigned long long prev = atomicCAS(&keys[slot], EDGE_EMPTY, key);
        if (prev == EDGE_EMPTY || prev == key) break;
        slot = (slot + 1) & mask;
    }
    atomicMin(&firstSide[slot], h);
    slotOf[h] = (int)slot;
}

// step 2, per triangle: which of its sides create a vertex, and how many triangles it becomes (1 + split sides).
__global__ void k_sub_count(const int* __restrict__ slotOf, const int* __restrict__ firstSide, int T, int* __restrict__ creates, int* __restrict__ pieces) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int split = 0;
    for (int j = 0; j < 3; j++) {
        int s = slotOf[3 * t + j];
        creates[3 * t + j] = (s >= 0 && firstSide[s] == 3 * t + j) ? 1 : 0;
        split += s >= 0;
    }
    pieces[t] = 1 + split;
}

struct SubTri {
    int v[3];
    V2 uv[3];
};
__device__ __forceinline__ void sub_emit(int* __restrict__ tri, double* __restrict__ uv, int at, int a, int b, int c, V2 ua, V2 ub, V2 uc) {
    tri[3 * (size_t)at] = a, tri[3 * (size_t)at + 1] = b, tri[3 * (size_t)at + 2] = c;
    double* o = uv + 6 * (size_t)at;
    o[0] = ua.x, o[1] = ua.y, o[2] = ub.x, o[3] = ub.y, o[4] = uc.x, o[5] = uc.y;
}

// step 3, per triangle: the midpoint vertices its sides create ((a + b) / 2 in single precision) and its 1-4 pieces, in
// the reference's order and orientation (MeshFlow.inl:177-217).
__global__ void k_sub_emit(const float* __restrict__ xyzIn, const int* __restrict__ triIn, const double* __restrict__ uvIn, int T, int V, const int* __restrict__ slotOf,
                           const int* __restrict__ firstSide, const int* __restrict__ createRank, const int* __restrict__ pieceOffset, float* __restrict__ xyzOut,
                           int* __restrict__ triOut, double* __restrict__ uvOut) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    SubTri s;
    for (int j = 0; j < 3; j++) s.v[j] = triIn[3 * t + j], s.uv[j] = v2(uvIn[6 * (size_t)t + 2 * j], uvIn[6 * (size_t)t + 2 * j + 1]);
    int e[3], split = 0;
    V2 mid[3];
    for (int j = 0; j < 3; j++) {
        int slot = slotOf[3 * t + j];
        e[j] = -1, mid[j] = v2(0, 0);
        if (slot < 0) continue;
        int creator = firstSide[slot];
        e[j] = V + createRank[creator];
        int j1 = (j + 1) % 3;
        mid[j] = v2(dadd(s.uv[j].x, s.uv[j1].x) / 2, dadd(s.uv[j].y, s.uv[j1].y) / 2);
        split++;
        if (creator == 3 * t + j)
            for (int k = 0; k < 3; k++) xyzOut[3 * (size_t)e[j] + k] = (xyzIn[3 * (size_t)s.v[j] + k] + xyzIn[3 * (size_t)s.v[j1] + k]) / 2;
    }
    int at = pieceOffset[t];
    if (split == 0) sub_emit(triOut, uvOut, at, s.v[0], s.v[1], s.v[2], s.uv[0], s.uv[1], s.uv[2]);
    else if (split == 1) {
        for (int j = 0; j < 3; j++)
            if (e[j] != -1) {
                int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                sub_emit(triOut, uvOut, at, s.v[j], e[j], s.v[j2], s.uv[j], mid[j], s.uv[j2]);
                sub_emit(triOut, uvOut, at + 1, s.v[j1], s.v[j2], e[j], s.uv[j1], s.uv[j2], mid[j]);
            }
    } else if (split == 2) {
        for (int j = 0; j < 3; j++)
            if (e[j] == -1) {
                int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                sub_emit(triOut, uvOut, at, e[j1], s.v[j2], e[j2], mid[j1], s.uv[j2], mid[j2]);
                sub_emit(triOut, uvOut, at + 1, s.v[j], s.v[j1], e[j2], s.uv[j], s.uv[j1], mid[j2]);
                sub_emit(triOut, uvOut, at + 2, s.v[j1], e[j1], e[j2], s.uv[j1], mid[j1], mid[j2]);
            }
    } else {
        for (int j = 0; j < 3; j++) sub_emit(triOut, uvOut, at + j, s.v[j], e[j], e[(j + 2) % 3], s.uv[j], mid[j], mid[(j + 2) % 3]);
        sub_emit(triOut, uvOut, at + 3, e[0], e[1], e[2], mid[0], mid[1], mid[2]);
    }
}

static unsigned pow2_at_least(unsigned long long x) {
    unsigned p = 1;
    while (p < x) p <<= 1;
    return p;
}

// Subdivide, MeshFlow.inl:223-232: sweeps until no side is longer than edgeLength. In and out: ctx->subXyz / subTri /
// subUv with ctx->subV vertices and ctx->subT triangles. Two integers come back to the host per sweep.
int subdivide_mesh(mof_ctx* ctx, double edgeLength, int* addedOut) {
    const int B = 256;
    int total = 0;
    DBuf<int> slotOf, firstSide, creates, createRank, pieces, pieceOffset, totals;
    DBuf<unsigned long long> keys;
    DBuf<float> xyzNext;
    DBuf<int> triNext;
    DBuf<double> uvNext;
    auto release = [&] {
        slotOf.release(), firstSide.release(), creates.release(), createRank.release(), pieces.release(), pieceOffset.release(), totals.release();
        keys.release(), xyzNext.release(), triNext.release(), uvNext.release();
    };
    int rc = [&]() -> int {
        MOF_CUDA(totals.alloc(2));
        for (int sweep = 0; sweep < 64; sweep++) {
            const int V = ctx->subV, T = ctx->subT, nSides = 3 * T;
            if ((long long)V + 3ll * T > INT_MAX / 4) return fail(ctx, MOF_E_INVALID, "mof_subdivide: the subdivided mesh would exceed 32-bit indexing");
            unsigned cap = pow2_at_least(2ull * nSides + 16);
            MOF_CUDA(keys.alloc(cap));
            MOF_CUDA(firstSide.alloc(cap));
            MOF_CUDA(slotOf.alloc(nSides));
            MOF_CUDA(creates.alloc(nSides));
            MOF_CUDA(createRank.alloc(nSides));
            MOF_CUDA(pieces.alloc(T));
            MOF_CUDA(pieceOffset.alloc(T));
            MOF_CUDA(cudaMemsetAsync(keys.p, 0xff, keys.bytes(), ctx->stream));
            MOF_CUDA(cudaMemsetAsync(firstSide.p, 0x7f, firstSide.bytes(), ctx->stream));
            MOF_LAUNCH(k_sub_mark, blocks_for(nSides, B), B, 0, ctx->subXyz.p, ctx->subTri.p, nSides, edgeLength * edgeLength, keys.p, firstSide.p, cap - 1, slotOf.p);
            MOF_LAUNCH(k_sub_count, blocks_for(T, B), B, 0, slotOf.p, firstSide.p, T, creates.p, pieces.p);
            MOF_TRY(exclusive_scan_int(ctx, creates.p, createRank.p, nSides, totals.p));
            MOF_TRY(exclusive_scan_int(ctx, pieces.p, pieceOffset.p, T, totals.p + 1));
            int h[2];
            MOF_CUDA(read_back(ctx, h, totals.p, 2));
            const int added = h[0], newT = h[1];
            if (!added) return MOF_OK;
            MOF_CUDA(xyzNext.alloc(3ull * (V + added)));
            MOF_CUDA(triNext.alloc(3ull * newT));
            MOF_CUDA(uvNext.alloc(6ull * newT));
            MOF_CUDA(cudaMemcpyAsync(xyzNext.p, ctx->subXyz.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToDevice, ctx->stream));
            MOF_LAUNCH(k_sub_emit, blocks_for(T, B), B, 0, ctx->subXyz.p, ctx->subTri.p, ctx->subUv.p, T, V, slotOf.p, firstSide.p, createRank.p, pieceOffset.p, xyzNext.p,
                       triNext.p, uvNext.p);
            std::swap(ctx->subXyz, xyzNext), std::swap(ctx->subTri, triNext), std::swap(ctx->subUv, uvNext);
            ctx->subV = V + added, ctx->subT = newT, total += added;
        }
        return fail(ctx, MOF_E_INVALID, "mof_subdivide: no convergence after 64 sweeps");
    }();
    release();
    if (addedOut) *addedOut = total;
    return rc;
}

}  // namespace mof
