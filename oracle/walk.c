/* TEST INFRASTRUCTURE ONLY — the CPU oracle's serial pieces, in plain C.
 *
 * A restatement (not a copy) of the reference's triangle-walk and texture-map routines, used by
 * oracle/mof_oracle.py through ctypes. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this; the product never does.
 *
 * Conventions shared with mof_oracle.py:
 *   - half-edge h = 3*t + j is the edge of triangle t opposite corner j (FEM.inl:597);
 *   - a 2x2 matrix is stored row-major in standard notation: m[0] m[1] / m[2] m[3];
 *   - the metric of triangle t is g[3t..3t+2] = (g00, g01, g11);
 *   - barycentric sample point (t, p) means corner0*(1-p0-p1) + corner1*p0 + corner2*p1.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const int* opp;      /* [3T] opposite half-edge or -1            (FEM.h:139) */
    const double* lin;   /* [3T][4] linear part of the edge transform (FEM.h:123) */
    const double* cst;   /* [3T][2] constant part                     (FEM.h:124) */
    const double* g;     /* [T][3] metric                             (FEM.h:148) */
} WalkMesh;

static double metric_dot(const double* g, const double* a, const double* b) {
    return a[0] * (g[0] * b[0] + g[1] * b[1]) + a[1] * (g[1] * b[0] + g[2] * b[1]);
}

/* RiemannianMesh::flow, FEM.inl:902-994. Moves (tIdx, p) for |flowTime| along the
 * piecewise-constant field vf (sign of flowTime picks the direction), re-reading the field every
 * minStepSize of arclength, unfolding across edges, stopping when the transported direction
 * opposes the local field. */
static void flow_point(const WalkMesh* m, const double* vf, double flowTime, int* tIdx, double* p, double minStepSize, double eps) {
    const int maxIters = 1000000;
    int inEdge = -1, t = *tIdx;
    double dir = flowTime < 0 ? -1. : 1.;
    double left = minStepSize;
    double v[2] = {vf[2 * t] * dir, vf[2 * t + 1] * dir};
    flowTime *= dir;
    for (int count = 0; count < maxIters; count++) {
        if (!(v[0] * v[0] + v[1] * v[1])) break;
        /* ray p + s v against the three sides: y=0 (edge 2), x=0 (edge 1), x+y=1 (edge 0) */
        double s = 0;
        int idx = -1;
        double cand[3] = {-p[1] / v[1], -p[0] / v[0], (1. - p[0] - p[1]) / (v[1] + v[0])};
        if (inEdge != 2 && cand[0] > 0) { double q = p[0] + v[0] * cand[0]; if (q >= -eps && q <= 1 + eps && cand[0] > s) idx = 2, s = cand[0]; }
        if (inEdge != 1 && cand[1] > 0) { double q = p[1] + v[1] * cand[1]; if (q >= -eps && q <= 1 + eps && cand[1] > s) idx = 1, s = cand[1]; }
        if (inEdge != 0 && cand[2] > 0) { double q = p[0] + v[0] * cand[2]; if (q >= -eps && q <= 1 + eps && cand[2] > s) idx = 0, s = cand[2]; }
        if (idx == -1) break;
        const double* gt = m->g + 3 * t;
        double vv = metric_dot(gt, v, v);
        double squareStep = vv * s * s;
        int refresh = 0;
        if (minStepSize > 0 && squareStep > left * left) { s = left / sqrt(vv); refresh = 1; }
        if (flowTime < s) { p[0] += v[0] * flowTime, p[1] += v[1] * flowTime; break; }
        if (refresh) {
            p[0] += v[0] * s, p[1] += v[1] * s, flowTime -= s;
            if (metric_dot(gt, v, vf + 2 * t) * dir < 0) break;
            v[0] = vf[2 * t] * dir, v[1] = vf[2 * t + 1] * dir;
            left = minStepSize;
            inEdge = -1;
        } else {
            p[0] += v[0] * s, p[1] += v[1] * s, flowTime -= s;
            int h = 3 * t + idx, o = m->opp[h];
            const double* L = m->lin + 4 * h;
            const double* c = m->cst + 2 * h;
            double q0 = L[0] * p[0] + L[1] * p[1] + c[0], q1 = L[2] * p[0] + L[3] * p[1] + c[1];
            double w0 = L[0] * v[0] + L[1] * v[1], w1 = L[2] * v[0] + L[3] * v[1];
            p[0] = q0, p[1] = q1, v[0] = w0, v[1] = w1;
            t = o / 3, inEdge = o % 3;
            left -= sqrt(squareStep);
        }
    }
    *tIdx = t;
}

void mof_oracle_flow(const int* opp, const double* lin, const double* cst, const double* g, const double* vf,
                     double flowTime, int* tIdx, double* p, double minStepSize, double eps) {
    WalkMesh m = {opp, lin, cst, g};
    flow_point(&m, vf, flowTime, tIdx, p, minStepSize, eps);
}

/* ResampleSignal, OpticalFlow.cpp:198-216 (+ Sample :180-186): one walk per triangle from its
 * centroid, barycentric sample of `in` at the landing point, scattered to the SOURCE triangle's
 * three vertices in triangle order and divided by the vertex valence. */
void mof_oracle_resample(int T, int V, const int* tri, const int* opp, const double* lin, const double* cst, const double* g,
                         const double* vf, const double* in, double* out, int C, double length) {
    WalkMesh m = {opp, lin, cst, g};
    int* counts = (int*)calloc((size_t)V, sizeof(int));
    memset(out, 0, sizeof(double) * (size_t)V * C);
    double c[16];
    for (int t = 0; t < T; t++) {
        int tt = t;
        double p[2] = {1. / 3, 1. / 3};
        flow_point(&m, vf, length, &tt, p, 1e-2, 0.);
        const int* corner = tri + 3 * tt;
        double w0 = 1. - p[0] - p[1];
        for (int k = 0; k < C; k++) c[k] = in[(size_t)corner[0] * C + k] * w0 + in[(size_t)corner[1] * C + k] * p[0] + in[(size_t)corner[2] * C + k] * p[1];
        for (int j = 0; j < 3; j++) {
            int vtx = tri[3 * t + j];
            for (int k = 0; k < C; k++) out[(size_t)vtx * C + k] += c[k];
            counts[vtx]++;
        }
    }
    for (int i = 0; i < V; i++) for (int k = 0; k < C; k++) out[(size_t)i * C + k] /= (double)counts[i];
    free(counts);
}

/* Sample of an RGB8 texture, MeshFlow.inl:66-84: v flipped, clamped to [0,1], bilinear. */
static void sample_texture(const unsigned char* tex, int W, int H, double u, double v, int bilinear, double* rgb) {
    v = 1 - v;
    u = fmin(1., fmax(0., u)), v = fmin(1., fmax(0., v));
    u *= W - 1, v *= H - 1;
    int x0 = (int)floor(u), y0 = (int)floor(v);
    if (bilinear) {
        double dx = u - x0, dy = v - y0;
        int x1 = x0 + 1 < W - 1 ? x0 + 1 : W - 1, y1 = y0 + 1 < H - 1 ? y0 + 1 : H - 1;
        for (int k = 0; k < 3; k++)
            rgb[k] = (double)tex[3 * (W * y0 + x0) + k] * ((1. - dx) * (1. - dy)) + (double)tex[3 * (W * y0 + x1) + k] * (dx * (1. - dy)) +
                     (double)tex[3 * (W * y1 + x1) + k] * (dx * dy) + (double)tex[3 * (W * y1 + x0) + k] * ((1. - dx) * dy);
    } else
        for (int k = 0; k < 3; k++) rgb[k] = (double)tex[3 * (W * y0 + x0) + k];
}

void mof_oracle_sample_texture(const unsigned char* tex, int W, int H, int n, const double* uv, int bilinear, double* rgb) {
    for (int i = 0; i < n; i++) sample_texture(tex, W, H, uv[2 * i], uv[2 * i + 1], bilinear, rgb + 3 * i);
}

/* InputTextureData::flow, OpticalFlow.cpp:501-515: per covered texel, walk its stored sample
 * point, map the landing point to uv through the triangle's texture coordinates, fetch. Texels
 * with srcT == -1 keep whatever `out` held. */
void mof_oracle_advect_texels(int W, int H, const int* srcT, const double* srcP, const int* opp, const double* lin, const double* cst,
                              const double* g, const double* vf, const double* triUV, const unsigned char* tex, double length, int bilinear, double* out) {
    WalkMesh m = {opp, lin, cst, g};
    for (int i = 0; i < W * H; i++) {
        if (srcT[i] == -1) continue;
        int t = srcT[i];
        double p[2] = {srcP[2 * i], srcP[2 * i + 1]};
        flow_point(&m, vf, length, &t, p, 1e-2, 0.);
        const double* uv = triUV + 6 * t;
        double w0 = 1. - p[0] - p[1];
        double qu = uv[0] * w0 + uv[2] * p[0] + uv[4] * p[1], qv = uv[1] * w0 + uv[3] * p[0] + uv[5] * p[1];
        sample_texture(tex, W, H, qu, qv, bilinear, out + 3 * i);
    }
}

/* InputTextureData::flow(frames), OpticalFlow.cpp:517-539: the sample points are carried along the flow in `frames - 1`
 * equal steps of 1 / (frames - 1) (minimum step 1e-2 * frames), each frame fetching the texture where the points stand;
 * frame 0 and the uncovered texels of every frame are the vertically flipped input. out: [frames][W*H][3]. */
void mof_oracle_advect_texels_frames(int W, int H, int frames, const int* srcT, const double* srcP, const int* opp, const double* lin, const double* cst,
                                     const double* g, const double* vf, const double* triUV, const unsigned char* tex, double sign, int bilinear, double* out) {
    WalkMesh m = {opp, lin, cst, g};
    const int n = W * H;
    const double length = sign / (frames - 1);
    for (int f = 0; f < frames; f++)
        for (int j = 0; j < H; j++)
            for (int i = 0; i < W; i++)
                for (int k = 0; k < 3; k++) out[3 * ((size_t)f * n + j * W + i) + k] = (double)tex[3 * ((H - j - 1) * W + i) + k];
    for (int i = 0; i < n; i++) {
        if (srcT[i] == -1) continue;
        int t = srcT[i];
        double p[2] = {srcP[2 * i], srcP[2 * i + 1]};
        for (int f = 1; f < frames; f++) {
            flow_point(&m, vf, length, &t, p, 1e-2 * frames, 0.);
            const double* uv = triUV + 6 * t;
            double w0 = 1. - p[0] - p[1];
            double qu = uv[0] * w0 + uv[2] * p[0] + uv[4] * p[1], qv = uv[1] * w0 + uv[3] * p[0] + uv[5] * p[1];
            sample_texture(tex, W, H, qu, qv, bilinear, out + 3 * ((size_t)f * n + i));
        }
    }
}

/* RiemannianMesh::exp, FEM.inl:835-899: straight line from p with initial velocity v (consumed as
 * it goes), unfolded across edges. Returns 0 on success, 1 if the ray misses the triangle (the
 * reference prints an error and exits), 2 on the iteration cap. */
static int exp_point(const WalkMesh* m, int* tIdx, double* p, double* v, double eps) {
    if (!(v[0] * v[0] + v[1] * v[1])) return 0;
    const int maxIters = 10000;
    int inEdge = -1, t = *tIdx;
    {
        int idx = -1;
        if (p[0] <= 0 && v[0] < 0) idx = 1;
        else if (p[1] <= 0 && v[1] < 0) idx = 2;
        else if (p[0] + p[1] >= 1 && v[0] + v[1] > 0) idx = 0;
        if (idx != -1) {
            int h = 3 * t + idx, o = m->opp[h];
            const double* L = m->lin + 4 * h;
            const double* c = m->cst + 2 * h;
            double q0 = L[0] * p[0] + L[1] * p[1] + c[0], q1 = L[2] * p[0] + L[3] * p[1] + c[1];
            double w0 = L[0] * v[0] + L[1] * v[1], w1 = L[2] * v[0] + L[3] * v[1];
            p[0] = q0, p[1] = q1, v[0] = w0, v[1] = w1;
            t = o / 3, inEdge = o % 3;
        }
    }
    for (int count = 0; count < maxIters; count++) {
        double s = 0;
        int idx = -1;
        double cand[3] = {-p[1] / v[1], -p[0] / v[0], (1. - p[0] - p[1]) / (v[1] + v[0])};
        if (inEdge != 2 && cand[0] > 0) { double q = p[0] + v[0] * cand[0]; if (q >= -eps && q <= 1 + eps && cand[0] > s) idx = 2, s = cand[0]; }
        if (inEdge != 1 && cand[1] > 0) { double q = p[1] + v[1] * cand[1]; if (q >= -eps && q <= 1 + eps && cand[1] > s) idx = 1, s = cand[1]; }
        if (inEdge != 0 && cand[2] > 0) { double q = p[0] + v[0] * cand[2]; if (q >= -eps && q <= 1 + eps && cand[2] > s) idx = 0, s = cand[2]; }
        if (idx == -1) { *tIdx = t; return 1; }
        if (s > 1) {
            p[0] += v[0], p[1] += v[1], v[0] = v[1] = 0;
            *tIdx = t;
            return 0;
        }
        p[0] += v[0] * s, p[1] += v[1] * s, v[0] -= v[0] * s, v[1] -= v[1] * s;
        int h = 3 * t + idx, o = m->opp[h];
        const double* L = m->lin + 4 * h;
        const double* c = m->cst + 2 * h;
        double q0 = L[0] * p[0] + L[1] * p[1] + c[0], q1 = L[2] * p[0] + L[3] * p[1] + c[1];
        double w0 = L[0] * v[0] + L[1] * v[1], w1 = L[2] * v[0] + L[3] * v[1];
        p[0] = q0, p[1] = q1, v[0] = w0, v[1] = w1;
        t = o / 3, inEdge = o % 3;
    }
    *tIdx = t;
    return 2;
}

/* BarycentricCoordinate, MeshFlow.inl:268-278. uv: three (u,v) corners. */
static void barycentric(const double* uv, double px, double py, double* b) {
    double w1x = uv[2] - uv[0], w1y = uv[3] - uv[1], w2x = uv[4] - uv[0], w2y = uv[5] - uv[1];
    double det = w1x * w2y - w2x * w1y, d = 1. / det;
    double rx = px - uv[0], ry = py - uv[1];
    b[0] = (w2y * rx - w2x * ry) * d;
    b[1] = (-w1y * rx + w1x * ry) * d;
}

/* RasterizeTriangle, MeshFlow.inl:281-337, including its first-writer-wins rule (:334: a later
 * triangle only replaces an earlier one when _b[1] >= 1, i.e. practically never). */
static void rasterize(const double* uvIn, int tIdx, int* srcT, double* srcP, int W, int H) {
    double v[6];
    for (int j = 0; j < 3; j++) v[2 * j] = uvIn[2 * j] * (W - 1), v[2 * j + 1] = uvIn[2 * j + 1] * (H - 1);
    int map[3];
    double y0 = uvIn[1], y1 = uvIn[3], y2 = uvIn[5];
    if (y0 <= y1 && y0 <= y2) { map[0] = 0; if (y1 <= y2) map[1] = 1, map[2] = 2; else map[1] = 2, map[2] = 1; }
    else if (y1 <= y0 && y1 <= y2) { map[0] = 1; if (y0 <= y2) map[1] = 0, map[2] = 2; else map[1] = 2, map[2] = 0; }
    else { map[0] = 2; if (y0 <= y1) map[1] = 0, map[2] = 1; else map[1] = 1, map[2] = 0; }
    double w[3][2];
    for (int j = 0; j < 3; j++) w[j][0] = v[2 * map[j]], w[j][1] = v[2 * map[j] + 1];
    int yStart = (int)ceil(w[0][1]), yEnd = (int)floor(w[2][1]);
    yStart = yStart < 0 ? 0 : (yStart > H - 1 ? H - 1 : yStart);
    yEnd = yEnd < 0 ? 0 : (yEnd > H - 1 ? H - 1 : yEnd);
    double src[2] = {w[0][0], w[0][1]}, sl0[2] = {w[1][0] - w[0][0], w[1][1] - w[0][1]}, sl1[2] = {w[2][0] - w[0][0], w[2][1] - w[0][1]};
    for (int y = yStart; y <= yEnd; y++) {
        if (y >= w[1][1]) {
            src[0] = w[2][0], src[1] = w[2][1];
            sl0[0] = w[1][0] - w[2][0], sl0[1] = w[1][1] - w[2][1];
            sl1[0] = w[0][0] - w[2][0], sl1[1] = w[0][1] - w[2][1];
        }
        if (sl0[1] == 0 || sl1[1] == 0) continue;
        double xa = src[0] + ((double)y - src[1]) * sl0[0] / sl0[1], xb = src[0] + ((double)y - src[1]) * sl1[0] / sl1[1];
        int xStart, xEnd;
        if (xa <= xb) xStart = (int)ceil(xa), xEnd = (int)floor(xb);
        else xStart = (int)ceil(xb), xEnd = (int)floor(xa);
        xStart = xStart < 0 ? 0 : (xStart > W - 1 ? W - 1 : xStart);
        xEnd = xEnd < 0 ? 0 : (xEnd > W - 1 ? W - 1 : xEnd);
        double b0[2], b1[2];
        barycentric(v, (double)xStart, (double)y, b0);
        barycentric(v, (double)xEnd, (double)y, b1);
        for (int x = xStart; x <= xEnd; x++) {
            double s = (double)(x - xStart) / (double)(xEnd - xStart);
            if (xStart == xEnd) s = 0.;
            double bx = b0[0] * (1. - s) + b1[0] * s, by = b0[1] * (1. - s) + b1[1] * s;
            int i = y * W + x;
            if (srcT[i] == -1 || (bx >= 0 && by >= 1 && bx + by <= 1)) srcT[i] = tIdx, srcP[2 * i] = bx, srcP[2 * i + 1] = by;
        }
    }
}

/* GetTextureSource, MeshFlow.inl:411-467: rasterise every triangle's uv footprint into a
 * texel -> (triangle, barycentric) map, grow it by padRadius rings of 4-neighbours, then pull
 * sample points that fell outside their triangle back onto the surface with exp (RemapSamplePoint,
 * :340-350). Returns the number of exp failures (0 expected). */
int mof_oracle_texture_source(int T, const double* triUV, int W, int H, int padRadius, const int* opp, const double* lin, const double* cst,
                              const double* g, int* srcT, double* srcP) {
    WalkMesh m = {opp, lin, cst, g};
    for (int i = 0; i < W * H; i++) srcT[i] = -1, srcP[2 * i] = srcP[2 * i + 1] = 0;
    for (int t = 0; t < T; t++) rasterize(triUV + 6 * t, t, srcT, srcP, W, H);
    int* update = (int*)malloc(sizeof(int) * (size_t)W * H);
    for (int r = 0; r < padRadius; r++) {
        for (int i = 0; i < W; i++) for (int j = 0; j < H; j++) {
            int idx = j * W + i;
            update[idx] = -1;
            if (srcT[idx] == -1) {
                for (int ii = -1; ii <= 1; ii++) if (i + ii >= 0 && i + ii < W && srcT[j * W + (i + ii)] != -1) update[idx] = srcT[j * W + (i + ii)];
                for (int jj = -1; jj <= 1; jj++) if (j + jj >= 0 && j + jj < H && srcT[(j + jj) * W + i] != -1) update[idx] = srcT[(j + jj) * W + i];
            }
        }
        for (int i = 0; i < W; i++) for (int j = 0; j < H; j++) {
            int idx = j * W + i, t = update[idx];
            if (t != -1) {
                srcT[idx] = t;
                barycentric(triUV + 6 * t, (double)i / (W - 1), (double)j / (H - 1), srcP + 2 * idx);
            }
        }
    }
    free(update);
    int failures = 0;
    for (int i = 0; i < W; i++) for (int j = 0; j < H; j++) {
        int idx = j * W + i;
        if (srcT[idx] == -1) continue;
        double* p = srcP + 2 * idx;
        if (p[0] >= 0 && p[1] >= 0 && p[0] + p[1] <= 1) continue;
        double v[2] = {p[0] - 1. / 3, p[1] - 1. / 3};
        p[0] = p[1] = 1. / 3;
        if (exp_point(&m, srcT + idx, p, v, 0.)) failures++;
    }
    return failures;
}

/* One pass of _Subdivide for textured triangles, MeshFlow.inl:158-220. New vertices are numbered
 * in the order their edges are first met (triangle by triangle, corner by corner). Positions are float (the reference's PlyVertex<float>), the length test squares the float
 * differences in double; uv is double.
 * Caller provides room: outV >= V + 3T vertices, outT >= 4T triangles. Returns the number of
 * vertices added; *nT receives the new triangle count. */
typedef struct { long long key; int idx; } EdgeSlot;

static int edge_lookup(EdgeSlot* table, size_t mask, long long key, int fresh) {
    size_t h = (size_t)((unsigned long long)key * 0x9E3779B97F4A7C15ull) & mask;
    while (table[h].idx != -1 && table[h].key != key) h = (h + 1) & mask;
    if (table[h].idx == -1) { table[h].key = key, table[h].idx = fresh; return -1; }
    return table[h].idx;
}

int mof_oracle_subdivide_pass(int V, int T, const float* vin, const int* tin, const double* uvin, double edgeLength,
                              float* vout, int* tout, double* uvout, int* nT) {
    size_t cap = 16;
    while (cap < (size_t)6 * T + 16) cap <<= 1;
    EdgeSlot* table = (EdgeSlot*)malloc(sizeof(EdgeSlot) * cap);
    for (size_t i = 0; i < cap; i++) table[i].idx = -1;
    memcpy(vout, vin, sizeof(float) * 3 * (size_t)V);
    int nv = V, nt = 0, added = 0;
#define EMIT(a, b, c, ua, ub, uc) do { tout[3 * nt] = (a), tout[3 * nt + 1] = (b), tout[3 * nt + 2] = (c); \
        uvout[6 * nt] = (ua)[0], uvout[6 * nt + 1] = (ua)[1], uvout[6 * nt + 2] = (ub)[0], uvout[6 * nt + 3] = (ub)[1], uvout[6 * nt + 4] = (uc)[0], uvout[6 * nt + 5] = (uc)[1]; nt++; } while (0)
    for (int i = 0; i < T; i++) {
        const int* tr = tin + 3 * i;
        const double* uv = uvin + 6 * i;
        int e[3] = {-1, -1, -1}, eCount = 0;
        double tex[3][2];
        for (int j = 0; j < 3; j++) {
            int i1 = tr[j], i2 = tr[(j + 1) % 3];
            float dx = vin[3 * i1] - vin[3 * i2], dy = vin[3 * i1 + 1] - vin[3 * i2 + 1], dz = vin[3 * i1 + 2] - vin[3 * i2 + 2];
            double l2 = (double)dx * (double)dx + (double)dy * (double)dy + (double)dz * (double)dz;
            if (l2 > edgeLength * edgeLength) {
                long long key = i1 > i2 ? (((long long)i1) << 32) | (long long)i2 : (((long long)i2) << 32) | (long long)i1;
                int idx = edge_lookup(table, cap - 1, key, nv);
                if (idx == -1) {
                    idx = nv;
                    for (int k = 0; k < 3; k++) vout[3 * nv + k] = (vin[3 * i1 + k] + vin[3 * i2 + k]) / 2;
                    nv++, added++;
                }
                e[j] = idx;
                tex[j][0] = (uv[2 * j] + uv[2 * ((j + 1) % 3)]) / 2, tex[j][1] = (uv[2 * j + 1] + uv[2 * ((j + 1) % 3) + 1]) / 2;
                eCount++;
            }
        }
#define UV(k) (uv + 2 * (k))
        if (eCount == 0) EMIT(tr[0], tr[1], tr[2], UV(0), UV(1), UV(2));
        else if (eCount == 1) {
            for (int j = 0; j < 3; j++) if (e[j] != -1) {
                EMIT(tr[j], e[j], tr[(j + 2) % 3], UV(j), tex[j], UV((j + 2) % 3));
                EMIT(tr[(j + 1) % 3], tr[(j + 2) % 3], e[j], UV((j + 1) % 3), UV((j + 2) % 3), tex[j]);
            }
        } else if (eCount == 2) {
            for (int j = 0; j < 3; j++) if (e[j] == -1) {
                EMIT(e[(j + 1) % 3], tr[(j + 2) % 3], e[(j + 2) % 3], tex[(j + 1) % 3], UV((j + 2) % 3), tex[(j + 2) % 3]);
                EMIT(tr[j], tr[(j + 1) % 3], e[(j + 2) % 3], UV(j), UV((j + 1) % 3), tex[(j + 2) % 3]);
                EMIT(tr[(j + 1) % 3], e[(j + 1) % 3], e[(j + 2) % 3], UV((j + 1) % 3), tex[(j + 1) % 3], tex[(j + 2) % 3]);
            }
        } else {
            for (int j = 0; j < 3; j++) EMIT(tr[j], e[j], e[(j + 2) % 3], UV(j), tex[j], tex[(j + 2) % 3]);
            EMIT(e[0], e[1], e[2], tex[0], tex[1], tex[2]);
        }
#undef UV
    }
#undef EMIT
    free(table);
    *nT = nt;
    return added;
}
