// Mesh-operator assembly on the GPU, straight from the face list.
//
// Everything here is built by gather: each output element is computed by one thread from the
// elements it depends on, in a fixed order, so results are deterministic run to run (the reference
// uses `omp atomic` sums, FEM.inl:1530, and unordered_map iteration order, SparseMatrix.inl:372-388).
// Sparsity patterns are stored with ascending columns in every row.
//
// Index conventions (as in the reference): half-edge h = 3t + j is the edge of triangle t opposite
// corner j, running corner (j+1)%3 -> corner (j+2)%3 (FEM.inl:597); a 2x2 matrix is row-major; the
// metric is (g00, g01, g11).
#include "mof_internal.cuh"

namespace mof {

// ------------------------------------------------------------------------------------- primitives

constexpr int SCAN_T = 256, SCAN_E = 4, SCAN_TILE = SCAN_T * SCAN_E;

__global__ void k_scan_tile(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ tileSum) {
    __shared__ int warpTot[SCAN_T / 32];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_E;
    int v[SCAN_E], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_E; k++) {
        int idx = base + k;
        int t = idx < n ? in[idx] : 0;
        v[k] = sum;
        sum += t;
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warpTot[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = lane < SCAN_T / 32 ? warpTot[lane] : 0, i2 = t;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, i2, o);
            if (lane >= o) i2 += y;
        }
        if (lane < SCAN_T / 32) warpTot[lane] = i2 - t;
        if (lane == SCAN_T / 32 - 1) tileSum[blockIdx.x] = i2;
    }
    __syncthreads();
    int off = warpTot[w] + inc - sum;
#pragma unroll
    for (int k = 0; k < SCAN_E; k++) {
        int idx = base + k;
        if (idx < n) out[idx] = v[k] + off;
    }
}

__global__ void k_scan_add(int* __restrict__ out, int n, const int* __restrict__ tileOff) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tileOff[i / SCAN_TILE];
}

__global__ void k_scan_total(const int* in, const int* out, int n, int* total) { total[0] = out[n - 1] + in[n - 1]; }

int exclusive_scan_int(mof_ctx* ctx, const int* in, int* out, int n, int* total) {
    if (n <= 0) return MOF_OK;
    int tiles = blocks_for(n, SCAN_TILE);
    DBuf<int> tileSum, tileOff;
    MOF_CUDA(tileSum.alloc(tiles));
    MOF_LAUNCH(k_scan_tile, tiles, SCAN_T, 0, in, out, n, tileSum.p);
    if (tiles > 1) {
        MOF_CUDA(tileOff.alloc(tiles));
        int rc = exclusive_scan_int(ctx, tileSum.p, tileOff.p, tiles, nullptr);
        if (rc != MOF_OK) { tileSum.release(), tileOff.release(); return rc; }
        MOF_LAUNCH(k_scan_add, blocks_for(n, 256), 256, 0, out, n, tileOff.p);
    }
    if (total) MOF_LAUNCH(k_scan_total, 1, 1, 0, in, out, n, total);
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    tileSum.release(), tileOff.release();
    return MOF_OK;
}

constexpr int RED_BLOCKS = kSMs * 4, RED_T = 256;

__global__ void k_reduce_partial(const double* __restrict__ in, long long n, double* __restrict__ partial) {
    __shared__ double sh[RED_T];
    double s = 0;
    for (long long i = (long long)blockIdx.x * RED_T + threadIdx.x; i < n; i += (long long)gridDim.x * RED_T) s += in[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = RED_T / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void k_reduce_final(const double* __restrict__ partial, int np, double* __restrict__ out) {
    __shared__ double sh[RED_T];
    double s = 0;
    for (int i = threadIdx.x; i < np; i += RED_T) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = RED_T / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// Deterministic sum: fixed grid, fixed tree.
int reduce_sum(mof_ctx* ctx, const double* in, long long n, double* out) {
    MOF_CUDA(ctx->dtmp1.reserve(RED_BLOCKS > 2048 ? RED_BLOCKS : 2048));
    MOF_LAUNCH(k_reduce_partial, RED_BLOCKS, RED_T, 0, in, n, ctx->dtmp1.p);
    MOF_LAUNCH(k_reduce_final, 1, RED_T, 0, ctx->dtmp1.p, RED_BLOCKS, out);
    return MOF_OK;
}

__global__ void k_fill_int(int* a, long long n, int v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

// ------------------------------------------------------------------------------ metric (a1, a2)

__device__ __forceinline__ double det3(const double* g) { return g[0] * g[2] - g[1] * g[1]; }
__device__ __forceinline__ void inv3(const double* g, double* gi) {
    double d = 1. / det3(g);
    gi[0] = g[2] * d, gi[1] = -g[1] * d, gi[2] = g[0] * d;
}
__device__ __forceinline__ void gmul(const double* g, double x, double y, double& ox, double& oy) {
    ox = g[0] * x + g[1] * y, oy = g[1] * x + g[2] * y;
}

// setMetricFromEmbedding, FEM.inl:1305-1323. flags[2]: vertex index out of range.
__global__ void k_metric(const double* __restrict__ pos, const int* __restrict__ tri, int T, int V, double* __restrict__ g,
                         double* __restrict__ sqdet, int* __restrict__ flags) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
    if ((unsigned)a >= (unsigned)V || (unsigned)b >= (unsigned)V || (unsigned)c >= (unsigned)V) {
        flags[2] = 1;
        g[3 * t] = g[3 * t + 2] = 1, g[3 * t + 1] = 0, sqdet[t] = 0;
        return;
    }
    double e0[3], e1[3];
    for (int k = 0; k < 3; k++) e0[k] = pos[3 * b + k] - pos[3 * a + k], e1[k] = pos[3 * c + k] - pos[3 * a + k];
    double g00 = e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2];
    double g01 = e0[0] * e1[0] + e0[1] * e1[1] + e0[2] * e1[2];
    double g11 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2];
    g[3 * t] = g00, g[3 * t + 1] = g01, g[3 * t + 2] = g11;
    double d = g00 * g11 - g01 * g01;
    if (!(d > 0)) flags[3] = 1;  // "[WARNING] Vanishing metric tensor determinant", FEM.inl:1317
    sqdet[t] = sqrt(d);
}

// makeUnitArea, FEM.inl:1283-1291, and area(i), :1301.
__global__ void k_metric_scale(double* __restrict__ g, double* __restrict__ area, int T, const double* __restrict__ sumSqdet) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double s = 2. / sumSqdet[0];
    double a = g[3 * t] * s, b = g[3 * t + 1] * s, c = g[3 * t + 2] * s;
    g[3 * t] = a, g[3 * t + 1] = b, g[3 * t + 2] = c;
    area[t] = sqrt(a * c - b * b) / 2.;
}

// --------------------------------------------------------------------- half-edge adjacency (a3)

__device__ __forceinline__ unsigned long long he_key(int a, int b) { return ((unsigned long long)(unsigned)a << 32) | (unsigned)b; }
__device__ __forceinline__ unsigned he_slot(unsigned long long k, unsigned mask) { return (unsigned)((k * 0x9E3779B97F4A7C15ull) >> 32) & mask; }
constexpr unsigned long long HE_EMPTY = ~0ull;

// setEdgeXForms first loop, FEM.inl:595-601. flags[0]: "[ERROR] Edge is occupied".
__global__ void k_he_insert(const int* __restrict__ tri, int nH, unsigned long long* keys, int* vals, unsigned mask, int* flags) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nH) return;
    int t = h / 3, j = h - 3 * t;
    unsigned long long key = he_key(tri[3 * t + (j + 1) % 3], tri[3 * t + (j + 2) % 3]);
    unsigned slot = he_slot(key, mask);
    while (true) {
        unsigned long long prev = atomicCAS(&keys[slot], HE_EMPTY, key);
        if (prev == HE_EMPTY) { vals[slot] = h; return; }
        if (prev == key) { flags[0] = 1; return; }
        slot = (slot + 1) & mask;
    }
}

// setEdgeXForms second loop, FEM.inl:603-613. flags[1]: a boundary half-edge exists.
__global__ void k_he_lookup(const int* __restrict__ tri, int nH, const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                            unsigned mask, int* __restrict__ opp, int* flags) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nH) return;
    int t = h / 3, j = h - 3 * t;
    unsigned long long key = he_key(tri[3 * t + (j + 2) % 3], tri[3 * t + (j + 1) % 3]);
    unsigned slot = he_slot(key, mask);
    while (true) {
        unsigned long long k = keys[slot];
        if (k == key) { opp[h] = vals[slot]; return; }
        if (k == HE_EMPTY) { opp[h] = -1; flags[1] = 1; return; }
        slot = (slot + 1) & mask;
    }
}

// FEM::Rotate90, FEM.inl:18-24.
__device__ __forceinline__ void rotate90(const double* g, const double* gi, double vx, double vy, double& wx, double& wy) {
    gmul(gi, -vy, vx, wx, wy);
    double tx, ty;
    gmul(g, vx, vy, tx, ty);
    double vn = tx * vx + ty * vy;
    gmul(g, wx, wy, tx, ty);
    double wn = tx * wx + ty * wy;
    if (wn) {
        double s = sqrt(vn / wn);
        wx *= s, wy *= s;
    }
}

// _setEdgeXForm, FEM.inl:550-590: the affine map of barycentric coordinates across half-edge h.
__global__ void k_edge_xforms(const double* __restrict__ g, const int* __restrict__ opp, int nH, double* __restrict__ xlin, double* __restrict__ xcst) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nH) return;
    int o = opp[h];
    if (o < 0) {
        xlin[4 * h] = 1, xlin[4 * h + 1] = 0, xlin[4 * h + 2] = 0, xlin[4 * h + 3] = 1;
        xcst[2 * h] = xcst[2 * h + 1] = 0;
        return;
    }
    const double cx[3] = {0., 1., 0.}, cy[3] = {0., 0., 1.};
    int t = h / 3, ot = o / 3;
    int v0 = (h + 1) % 3, v1 = (h + 2) % 3, ov0 = (o + 1) % 3, ov1 = (o + 2) % 3;
    double gt[3] = {g[3 * t], g[3 * t + 1], g[3 * t + 2]}, go[3] = {g[3 * ot], g[3 * ot + 1], g[3 * ot + 2]};
    double gti[3], goi[3];
    inv3(gt, gti), inv3(go, goi);
    double ex = cx[v1] - cx[v0], ey = cy[v1] - cy[v0];
    double ox = -(cx[ov1] - cx[ov0]), oy = -(cy[ov1] - cy[ov0]);
    double tx, ty;
    gmul(gt, ex, ey, tx, ty);
    double len = sqrt(ex * tx + ey * ty);
    ex /= len, ey /= len;
    gmul(go, ox, oy, tx, ty);
    len = sqrt(ox * tx + oy * ty);
    ox /= len, oy /= len;
    double px, py, qx, qy;
    rotate90(gt, gti, ex, ey, px, py);
    rotate90(go, goi, ox, oy, qx, qy);
    // M = [e p] (columns), oM = [o q]; linear = oM * M^-1
    double d = 1. / (ex * py - px * ey);
    double i00 = py * d, i01 = -px * d, i10 = -ey * d, i11 = ex * d;
    double l00 = ox * i00 + qx * i10, l01 = ox * i01 + qx * i11, l10 = oy * i00 + qy * i10, l11 = oy * i01 + qy * i11;
    double sx = cx[v0] + cx[v1], sy = cy[v0] + cy[v1], osx = cx[ov0] + cx[ov1], osy = cy[ov0] + cy[ov1];
    xlin[4 * h] = l00, xlin[4 * h + 1] = l01, xlin[4 * h + 2] = l10, xlin[4 * h + 3] = l11;
    xcst[2 * h] = (osx - (l00 * sx + l01 * sy)) / 2.;
    xcst[2 * h + 1] = (osy - (l10 * sx + l11 * sy)) / 2.;
}

// ------------------------------------------------------ scalar mass / stiffness CSR (a4), V x V

__global__ void k_count_outgoing(const int* __restrict__ tri, int nH, int* cnt) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nH) return;
    int t = h / 3, j = h - 3 * t;
    atomicAdd(&cnt[tri[3 * t + (j + 1) % 3]], 1);
}

// One (column, source half-edge) pair per outgoing half-edge plus the diagonal (he = -1).
__global__ void k_scalar_fill(const int* __restrict__ tri, int nH, int V, int* cursor, int* __restrict__ col, int* __restrict__ he) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nH) {
        int t = i / 3, j = i - 3 * t;
        int a = tri[3 * t + (j + 1) % 3], b = tri[3 * t + (j + 2) % 3];
        int slot = atomicAdd(&cursor[a], 1);
        col[slot] = b, he[slot] = i;
    } else if (i < nH + V) {
        int a = i - nH;
        int slot = atomicAdd(&cursor[a], 1);
        col[slot] = a, he[slot] = -1;
    }
}

// Ascending columns in each row (short rows: insertion sort by one thread), payload carried along.
__global__ void k_sort_rows(const int* __restrict__ rowptr, int rows, int* __restrict__ col, int* __restrict__ payload) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int b = rowptr[r], e = rowptr[r + 1];
    for (int i = b + 1; i < e; i++) {
        int c = col[i], p = payload ? payload[i] : 0, k = i - 1;
        while (k >= b && col[k] > c) {
            col[k + 1] = col[k];
            if (payload) payload[k + 1] = payload[k];
            k--;
        }
        col[k + 1] = c;
        if (payload) payload[k + 1] = p;
    }
}

// Local stiffness entry (i,j) of a triangle: grad_i . ginv grad_j / 2 * sqrt det, FEM.inl:480-496.
__device__ __forceinline__ double local_stiffness(const double* gi, double sq, int i, int j) {
    const double gx[3] = {-1., 1., 0.}, gy[3] = {-1., 0., 1.};
    double tx, ty;
    gmul(gi, gx[j], gy[j], tx, ty);
    return (gx[i] * tx + gy[i] * ty) / 2. * sq;
}

// _scalarMatrix, FEM.inl:1507-1547 with SetScalarMassMatrix / SetScalarStiffnessMatrix (:439-496): entry (a,b)
// gathers the local (corner a, corner b) entries of the two triangles on edge ab; the diagonal
// gathers the fan of a in column order. flags[4]: a vertex with no triangle.
__global__ void k_scalar_values(const int* __restrict__ rowptr, const int* __restrict__ he, const int* __restrict__ opp, const double* __restrict__ g,
                                int V, double* __restrict__ mass, double* __restrict__ stiff, int* flags) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= V) return;
    int b0 = rowptr[a], e0 = rowptr[a + 1];
    if (e0 - b0 < 2) flags[4] = 1;
    double dm = 0, ds = 0;
    int diag = -1;
    for (int k = b0; k < e0; k++) {
        int h = he[k];
        if (h < 0) { diag = k; continue; }
        int t = h / 3, j = h - 3 * t, ia = (j + 1) % 3, ib = (j + 2) % 3;
        double gt[3] = {g[3 * t], g[3 * t + 1], g[3 * t + 2]}, gi[3];
        double sq = sqrt(det3(gt));
        inv3(gt, gi);
        double m = sq * (1. / 24), s = local_stiffness(gi, sq, ia, ib);
        dm += sq * (1. / 12), ds += local_stiffness(gi, sq, ia, ia);
        int o = opp[h];
        if (o >= 0) {
            int t2 = o / 3, j2 = o - 3 * t2, ia2 = (j2 + 2) % 3, ib2 = (j2 + 1) % 3;
            double g2[3] = {g[3 * t2], g[3 * t2 + 1], g[3 * t2 + 2]}, gi2[3];
            double sq2 = sqrt(det3(g2));
            inv3(g2, gi2);
            m += sq2 * (1. / 24), s += local_stiffness(gi2, sq2, ia2, ib2);
        }
        mass[k] = m, stiff[k] = s;
    }
    if (diag >= 0) mass[diag] = dm, stiff[diag] = ds;
}

// ------------------------------------------------------------- Whitney numbering, P (a6, a7)

// InitializeCoefficients, Whitney.inl:39-51: a dof per undirected edge, numbered by the rank of its
// first half-edge in (t,j) order; the other half-edge carries the opposite orientation.
__global__ void k_first_flags(const int* __restrict__ opp, int nH, int* __restrict__ flag) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h > nH) return;
    flag[h] = h < nH && (opp[h] < 0 || h < opp[h]) ? 1 : 0;
}
__global__ void k_numbering(const int* __restrict__ opp, const int* __restrict__ flag, const int* __restrict__ rank, int nH,
                            int* __restrict__ reduced, int* __restrict__ expanded, int* __restrict__ positive) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nH) return;
    if (flag[h]) reduced[h] = rank[h], expanded[rank[h]] = h, positive[h] = 1;
    else reduced[h] = rank[opp[h]], positive[h] = 0;
}

// InitializeProlonagtionOperator, Whitney.inl:65-88: P[t][k] = +-ginv (grad[k+2]-grad[k+1]) / 3.
__global__ void k_prolongation(const double* __restrict__ g, const int* __restrict__ positive, int T, double* __restrict__ P) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const double gx[3] = {-1., 1., 0.}, gy[3] = {-1., 0., 1.};
    double gt[3] = {g[3 * t], g[3 * t + 1], g[3 * t + 2]}, gi[3];
    inv3(gt, gi);
    for (int k = 0; k < 3; k++) {
        double dx = (gx[(k + 2) % 3] - gx[(k + 1) % 3]) / 3.0, dy = (gy[(k + 2) % 3] - gy[(k + 1) % 3]) / 3.0;
        double px, py;
        gmul(gi, dx, dy, px, py);
        if (!positive[3 * t + k]) px *= -1, py *= -1;
        P[6 * t + 2 * k] = px, P[6 * t + 2 * k + 1] = py;
    }
}

// ------------------------------------------------------------- Whitney smooth operator (a8), E x E

// m0: barycentric vertex areas, Whitney.inl:122-126.
__global__ void k_vertex_area(const int* __restrict__ rowptr, const int* __restrict__ he, const double* __restrict__ area, int V, double* __restrict__ m0) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= V) return;
    double s = 0;
    for (int k = rowptr[a]; k < rowptr[a + 1]; k++)
        if (he[k] >= 0) s += area[he[k] / 3] / 3.;
    m0[a] = s;
}

__device__ __forceinline__ double cot_term(const double* __restrict__ g, const double* __restrict__ area, int h) {
    const double gx[3] = {-1., 1., 0.}, gy[3] = {-1., 0., 1.};
    int t = h / 3, v = h - 3 * t;
    double gt[3] = {g[3 * t], g[3 * t + 1], g[3 * t + 2]}, gi[3];
    inv3(gt, gi);
    double tx, ty;
    gmul(gi, gx[(v + 2) % 3], gy[(v + 2) % 3], tx, ty);
    return -area[t] * (gx[(v + 1) % 3] * tx + gy[(v + 1) % 3] * ty);
}

// m1: cotangent edge weights, Whitney.inl:142-160.
__global__ void k_edge_weight(const double* __restrict__ g, const double* __restrict__ area, const int* __restrict__ expanded,
                              const int* __restrict__ opp, int E, double* __restrict__ m1) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int h = expanded[e], o = opp[h];
    double r = cot_term(g, area, h);
    if (o >= 0) r += cot_term(g, area, o);
    m1[e] = r;
}

// Row e = (a->b) holds every edge touching a or b: deg(a) + deg(b) - 1 entries.
__global__ void k_whitney_rowsize(const int* __restrict__ tri, const int* __restrict__ expanded, const int* __restrict__ sRowptr, int E, int* __restrict__ size) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > E) return;
    if (e == E) { size[e] = 0; return; }
    int h = expanded[e], t = h / 3, j = h - 3 * t;
    int a = tri[3 * t + (j + 1) % 3], b = tri[3 * t + (j + 2) % 3];
    size[e] = (sRowptr[a + 1] - sRowptr[a] - 1) + (sRowptr[b + 1] - sRowptr[b] - 1) - 1;
}

// The E x E operators are stored SLICED (SELL-32, mof_internal.cuh): rows in groups of 32, each group
// padded to its longest row and stored entry-major, so that a warp working on 32 consecutive rows
// reads 32 consecutive words per entry index. size32[s] = 32 * (longest row of slice s).
__global__ void k_slice_sizes(const int* __restrict__ rowptr, int n, int slices, int* __restrict__ size32) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > slices) return;
    int longest = 0;
    if (s < slices)
        for (int r = 32 * s; r < min(n, 32 * s + 32); r++) longest = max(longest, rowptr[r + 1] - rowptr[r]);
    size32[s] = 32 * longest;
}

// One thread per (padded) row: the real entries, then the padding (column = the row itself, value 0).
__global__ void k_whitney_fill(const int* __restrict__ tri, const int* __restrict__ expanded, const int* __restrict__ reduced, const int* __restrict__ opp,
                               const int* __restrict__ sRowptr, const int* __restrict__ sCol, const int* __restrict__ sHe, const int* __restrict__ sliceBase, int E,
                               int slices, int* __restrict__ wCol) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 32 * slices) return;
    const int longest = (sliceBase[(e >> 5) + 1] - sliceBase[e >> 5]) >> 5;
    int j = 0;
    if (e < E) {
        int h = expanded[e], t = h / 3, c = h - 3 * t;
        int a = tri[3 * t + (c + 1) % 3], b = tri[3 * t + (c + 2) % 3];
        for (int k = sRowptr[a]; k < sRowptr[a + 1]; k++)
            if (sHe[k] >= 0) wCol[sell_pos(sliceBase, e, j++)] = reduced[sHe[k]];
        for (int k = sRowptr[b]; k < sRowptr[b + 1]; k++)
            if (sHe[k] >= 0 && sCol[k] != a) wCol[sell_pos(sliceBase, e, j++)] = reduced[sHe[k]];
    }
    for (; j < longest; j++) wCol[sell_pos(sliceBase, e, j)] = e < E ? e : 0;
}

// Ascending columns within each sliced row (insertion sort over the row's strided entries).
__global__ void k_sort_rows_sell(const int* __restrict__ rowptr, const int* __restrict__ sliceBase, int rows, int* __restrict__ col) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int len = rowptr[r + 1] - rowptr[r];
    size_t base = sell_pos(sliceBase, r, 0);
    for (int i = 1; i < len; i++) {
        int c = col[base + 32 * (size_t)i], k = i - 1;
        while (k >= 0 && col[base + 32 * (size_t)k] > c) {
            col[base + 32 * (size_t)(k + 1)] = col[base + 32 * (size_t)k];
            k--;
        }
        col[base + 32 * (size_t)(k + 1)] = c;
    }
}

// InitializeSmoothOperator, Whitney.inl:92-180: S = (d1^T m2 d1 + m1 d0 m0^-1 d0^T m1) / 2, entry by entry.
__global__ void k_whitney_values(const int* __restrict__ tri, const int* __restrict__ expanded, const int* __restrict__ reduced, const int* __restrict__ positive,
                                 const int* __restrict__ opp, const double* __restrict__ area, const double* __restrict__ m0, const double* __restrict__ m1,
                                 const int* __restrict__ wRowptr, const int* __restrict__ sliceBase, const int* __restrict__ wCol, int E,
                                 double* __restrict__ wS) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int h = expanded[e], t1 = h / 3, j1 = h - 3 * t1, o = opp[h];
    int a = tri[3 * t1 + (j1 + 1) % 3], b = tri[3 * t1 + (j1 + 2) % 3];
    int t2 = o >= 0 ? o / 3 : -1;
    double w1 = 1. / area[t1], w2 = t2 >= 0 ? 1. / area[t2] : 0.;
    double se1 = positive[h] ? 1. : -1., se2 = o >= 0 ? (positive[o] ? 1. : -1.) : 0.;
    double me = m1[e], ia = 1.0 / m0[a], ib = 1.0 / m0[b];
    const int len = wRowptr[e + 1] - wRowptr[e];
    for (int jj = 0; jj < len; jj++) {
        const size_t k = sell_pos(sliceBase, e, jj);
        int f = wCol[k];
        double rot = 0;
        for (int q = 0; q < 3; q++) {
            if (reduced[3 * t1 + q] == f) rot += (se1 * w1) * (positive[3 * t1 + q] ? 1. : -1.);
            if (t2 >= 0 && reduced[3 * t2 + q] == f) rot += (se2 * w2) * (positive[3 * t2 + q] ? 1. : -1.);
        }
        int hf = expanded[f], tf = hf / 3, jf = hf - 3 * tf;
        int fa = tri[3 * tf + (jf + 1) % 3], fb = tri[3 * tf + (jf + 2) % 3];
        double mf = m1[f], div = 0;
        // d0[e][tail] = -1, d0[e][head] = +1 (Whitney.inl:104-105)
        if (a == fa) div += (((me * -1.) * ia) * -1.) * mf;
        if (a == fb) div += (((me * -1.) * ia) * 1.) * mf;
        if (b == fa) div += (((me * 1.) * ib) * -1.) * mf;
        if (b == fb) div += (((me * 1.) * ib) * 1.) * mf;
        wS[k] = (rot + div) * 0.5;
    }
}

// ------------------------------------------------------------------------------------- driver

static unsigned next_pow2(unsigned long long x) {
    unsigned long long p = 1;
    while (p < x) p <<= 1;
    return (unsigned)p;
}

int build_mesh_operators(mof_ctx* ctx) {
    const int V = ctx->V, T = ctx->T, nH = 3 * T, B = 256;
    PhaseTimer pt(ctx);
    MOF_CUDA(ctx->g.alloc(3ull * T));
    MOF_CUDA(ctx->area.alloc(T));
    MOF_CUDA(ctx->opp.alloc(nH));
    MOF_CUDA(ctx->xlin.alloc(4ull * nH));
    MOF_CUDA(ctx->xcst.alloc(2ull * nH));
    MOF_CUDA(ctx->flags.alloc(16));
    MOF_CUDA(ctx->scalars.alloc(SC_COUNT));
    MOF_CUDA(cudaMemsetAsync(ctx->flags.p, 0, ctx->flags.bytes(), ctx->stream));
    MOF_CUDA(cudaMemsetAsync(ctx->scalars.p, 0, ctx->scalars.bytes(), ctx->stream));

    // a1, a2: metric, unit area
    MOF_CUDA(ctx->dtmp0.alloc(T));
    MOF_LAUNCH(k_metric, blocks_for(T, B), B, 0, ctx->pos.p, ctx->tri.p, T, V, ctx->g.p, ctx->dtmp0.p, ctx->flags.p);
    int hflags[16];
    MOF_CUDA(cudaMemcpyAsync(hflags, ctx->flags.p, sizeof(hflags), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hflags[2]) return fail(ctx, MOF_E_INVALID, "[ERROR] triangle refers to a vertex index outside [0,V)");
    MOF_TRY(reduce_sum(ctx, ctx->dtmp0.p, T, ctx->scalars.p + SC_AREA_SCALE));
    MOF_LAUNCH(k_metric_scale, blocks_for(T, B), B, 0, ctx->g.p, ctx->area.p, T, ctx->scalars.p + SC_AREA_SCALE);

    pt.mark("  metric, unit area");
    // a3: opposite half-edges through an open-addressing table, then the edge transforms
    unsigned cap = next_pow2(2ull * nH + 16);
    MOF_CUDA(ctx->hashKeys.alloc(cap));
    MOF_CUDA(ctx->itmp0.alloc(cap));
    MOF_CUDA(cudaMemsetAsync(ctx->hashKeys.p, 0xff, ctx->hashKeys.bytes(), ctx->stream));
    MOF_LAUNCH(k_he_insert, blocks_for(nH, B), B, 0, ctx->tri.p, nH, ctx->hashKeys.p, ctx->itmp0.p, cap - 1, ctx->flags.p);
    MOF_LAUNCH(k_he_lookup, blocks_for(nH, B), B, 0, ctx->tri.p, nH, ctx->hashKeys.p, ctx->itmp0.p, cap - 1, ctx->opp.p, ctx->flags.p);
    MOF_CUDA(cudaMemcpyAsync(hflags, ctx->flags.p, sizeof(hflags), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hflags[0]) return fail(ctx, MOF_E_MESH, "[ERROR] Edge is occupied");
    if (hflags[1]) return fail(ctx, MOF_E_MESH, "[ERROR] Boundary edge (TriangleMesh::unfold)");
    MOF_LAUNCH(k_edge_xforms, blocks_for(nH, B), B, 0, ctx->g.p, ctx->opp.p, nH, ctx->xlin.p, ctx->xcst.p);

    pt.mark("  half-edges, edge transforms");
    // a4: V x V pattern (diagonal + one entry per outgoing half-edge), sorted, then values
    MOF_CUDA(ctx->itmp1.alloc(V + 1));
    MOF_CUDA(ctx->sRowptr.alloc(V + 1));
    MOF_LAUNCH(k_fill_int, blocks_for(V + 1, B), B, 0, ctx->itmp1.p, (long long)V + 1, 1);
    MOF_CUDA(cudaMemsetAsync(ctx->itmp1.p + V, 0, sizeof(int), ctx->stream));
    MOF_LAUNCH(k_count_outgoing, blocks_for(nH, B), B, 0, ctx->tri.p, nH, ctx->itmp1.p);
    MOF_TRY(exclusive_scan_int(ctx, ctx->itmp1.p, ctx->sRowptr.p, V + 1, nullptr));
    int nnzS = 0;
    MOF_CUDA(read_back(ctx, &nnzS, ctx->sRowptr.p + V));
    ctx->nnzS = nnzS;
    MOF_CUDA(ctx->sCol.alloc(nnzS));
    MOF_CUDA(ctx->sHe.alloc(nnzS));
    MOF_CUDA(ctx->sMass.alloc(nnzS));
    MOF_CUDA(ctx->sStiff.alloc(nnzS));
    MOF_CUDA(ctx->sSys.alloc(nnzS));
    MOF_CUDA(ctx->sDinv.alloc(6ull * V));
    MOF_CUDA(cudaMemcpyAsync(ctx->itmp1.p, ctx->sRowptr.p, sizeof(int) * (V + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    MOF_LAUNCH(k_scalar_fill, blocks_for(nH + V, B), B, 0, ctx->tri.p, nH, V, ctx->itmp1.p, ctx->sCol.p, ctx->sHe.p);
    MOF_LAUNCH(k_sort_rows, blocks_for(V, B), B, 0, ctx->sRowptr.p, V, ctx->sCol.p, ctx->sHe.p);
    MOF_LAUNCH(k_scalar_values, blocks_for(V, B), B, 0, ctx->sRowptr.p, ctx->sHe.p, ctx->opp.p, ctx->g.p, V, ctx->sMass.p, ctx->sStiff.p, ctx->flags.p);

    {  // the scalar pattern once more, sliced (SELL-32): slice offsets and columns per mesh, values per system (multigrid.cu)
        MOF_TRY(csr_to_sell(ctx, V, ctx->sRowptr.p, ctx->sCol.p, ctx->sMass.p, ctx->sSliceBase, ctx->sColSell, ctx->sSysSell));
        ctx->sPadded = (long long)ctx->sColSell.n;
    }
    pt.mark("  scalar operators");
    // a6: Whitney dof numbering
    MOF_CUDA(ctx->itmp0.alloc(nH + 1));
    MOF_CUDA(ctx->itmp2.alloc(nH + 1));
    MOF_CUDA(ctx->reduced.alloc(nH));
    MOF_CUDA(ctx->positive.alloc(nH));
    MOF_LAUNCH(k_first_flags, blocks_for(nH + 1, B), B, 0, ctx->opp.p, nH, ctx->itmp0.p);
    MOF_TRY(exclusive_scan_int(ctx, ctx->itmp0.p, ctx->itmp2.p, nH + 1, nullptr));
    int E = 0;
    MOF_CUDA(read_back(ctx, &E, ctx->itmp2.p + nH));
    ctx->E = E;
    MOF_CUDA(ctx->expanded.alloc(E));
    MOF_LAUNCH(k_numbering, blocks_for(nH, B), B, 0, ctx->opp.p, ctx->itmp0.p, ctx->itmp2.p, nH, ctx->reduced.p, ctx->expanded.p, ctx->positive.p);

    // a7: prolongation
    MOF_CUDA(ctx->P.alloc(6ull * T));
    MOF_LAUNCH(k_prolongation, blocks_for(T, B), B, 0, ctx->g.p, ctx->positive.p, T, ctx->P.p);

    pt.mark("  numbering, prolongation");
    // a8: smooth operator
    MOF_CUDA(ctx->m0.alloc(V));
    MOF_CUDA(ctx->m1.alloc(E));
    MOF_LAUNCH(k_vertex_area, blocks_for(V, B), B, 0, ctx->sRowptr.p, ctx->sHe.p, ctx->area.p, V, ctx->m0.p);
    MOF_LAUNCH(k_edge_weight, blocks_for(E, B), B, 0, ctx->g.p, ctx->area.p, ctx->expanded.p, ctx->opp.p, E, ctx->m1.p);
    MOF_CUDA(ctx->itmp1.alloc(E + 1));
    MOF_CUDA(ctx->wRowptr.alloc(E + 1));
    MOF_LAUNCH(k_whitney_rowsize, blocks_for(E + 1, B), B, 0, ctx->tri.p, ctx->expanded.p, ctx->sRowptr.p, E, ctx->itmp1.p);
    MOF_TRY(exclusive_scan_int(ctx, ctx->itmp1.p, ctx->wRowptr.p, E + 1, nullptr));
    int nnzW = 0;
    MOF_CUDA(read_back(ctx, &nnzW, ctx->wRowptr.p + E));
    ctx->nnzW = nnzW;
    // sliced layout: slice sizes -> slice offsets -> padded entry count
    const int slices = (E + 31) / 32;
    ctx->wSlices = slices;
    MOF_CUDA(ctx->itmp0.alloc(slices + 1));
    MOF_CUDA(ctx->wSliceBase.alloc(slices + 1));
    MOF_LAUNCH(k_slice_sizes, blocks_for(slices + 1, B), B, 0, ctx->wRowptr.p, E, slices, ctx->itmp0.p);
    MOF_TRY(exclusive_scan_int(ctx, ctx->itmp0.p, ctx->wSliceBase.p, slices + 1, nullptr));
    int padded = 0;
    MOF_CUDA(read_back(ctx, &padded, ctx->wSliceBase.p + slices));
    if (padded < 0 || (long long)padded < nnzW)  // the padded entry count (a 32-bit scan) wrapped: the mesh is beyond what int offsets index
        return fail(ctx, MOF_E_INVALID, "mof_set_mesh: the padded Whitney pattern does not fit 32-bit offsets (mesh too large)");
    ctx->wPadded = padded;
    MOF_CUDA(ctx->wCol.alloc((size_t)padded));
    MOF_CUDA(ctx->wS.alloc((size_t)padded));
    MOF_CUDA(ctx->wA.alloc((size_t)padded));
    MOF_CUDA(ctx->wDinv.alloc(E));
    MOF_CUDA(cudaMemsetAsync(ctx->wS.p, 0, sizeof(double) * (size_t)padded, ctx->stream));  // padding entries stay 0 for good
    MOF_CUDA(cudaMemsetAsync(ctx->wA.p, 0, sizeof(double) * (size_t)padded, ctx->stream));
    MOF_LAUNCH(k_whitney_fill, blocks_for(32ll * slices, B), B, 0, ctx->tri.p, ctx->expanded.p, ctx->reduced.p, ctx->opp.p, ctx->sRowptr.p, ctx->sCol.p,
               ctx->sHe.p, ctx->wSliceBase.p, E, slices, ctx->wCol.p);
    MOF_LAUNCH(k_sort_rows_sell, blocks_for(E, B), B, 0, ctx->wRowptr.p, ctx->wSliceBase.p, E, ctx->wCol.p);
    MOF_LAUNCH(k_whitney_values, blocks_for(E, B), B, 0, ctx->tri.p, ctx->expanded.p, ctx->reduced.p, ctx->positive.p, ctx->opp.p, ctx->area.p, ctx->m0.p,
               ctx->m1.p, ctx->wRowptr.p, ctx->wSliceBase.p, ctx->wCol.p, E, ctx->wS.p);

    MOF_CUDA(cudaMemcpyAsync(hflags, ctx->flags.p, sizeof(hflags), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hflags[4]) return fail(ctx, MOF_E_MESH, "[ERROR] vertex without a triangle (singular mass matrix)");
    // (the scratch buffers stay allocated for the next mesh)

    pt.mark("  Whitney operator");
    // flow-side buffers
    MOF_CUDA(ctx->coeffs.alloc(E));
    MOF_CUDA(ctx->tfield.alloc(2ull * T));
    MOF_CUDA(ctx->fb.alloc(E));
    MOF_CUDA(ctx->fx.alloc(E));
    MOF_CUDA(ctx->dataD.alloc(3ull * T));
    MOF_CUDA(ctx->dataRhs.alloc(2ull * T));
    MOF_CUDA(ctx->tsample6.alloc(6ull * T));
    MOF_CUDA(cudaMemsetAsync(ctx->coeffs.p, 0, ctx->coeffs.bytes(), ctx->stream));
    MOF_CUDA(cudaMemsetAsync(ctx->tfield.p, 0, ctx->tfield.bytes(), ctx->stream));
    return MOF_OK;
}

// ------------------------------------------------------------------ CSR <-> sliced layout conversions

__global__ void k_csr_to_sell(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, const int* __restrict__ sliceBase, int n,
                              int slices, int* __restrict__ sCol, double* __restrict__ sVal) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= 32 * slices) return;
    const int longest = (sliceBase[(r >> 5) + 1] - sliceBase[r >> 5]) >> 5;
    int j = 0;
    if (r < n)
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++, j++) {
            size_t p = sell_pos(sliceBase, r, j);
            sCol[p] = col[k], sVal[p] = val[k];
        }
    for (; j < longest; j++) {
        size_t p = sell_pos(sliceBase, r, j);
        sCol[p] = r < n ? r : 0, sVal[p] = 0.;
    }
}

__global__ void k_sell_to_csr(const int* __restrict__ rowptr, const int* __restrict__ sliceBase, const int* __restrict__ sCol, const double* __restrict__ sVal, int n,
                              int* __restrict__ col, double* __restrict__ val) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int j = 0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; k++, j++) {
        size_t p = sell_pos(sliceBase, r, j);
        col[k] = sCol[p], val[k] = sVal[p];
    }
}

int csr_to_sell(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, DBuf<int>& sliceBase, DBuf<int>& sCol, DBuf<double>& sVal) {
    const int slices = (n + 31) / 32, B = 256;
    DBuf<int> sizes;
    MOF_CUDA(sizes.alloc(slices + 1));
    MOF_CUDA(sliceBase.alloc(slices + 1));
    MOF_LAUNCH(k_slice_sizes, blocks_for(slices + 1, B), B, 0, rowptr, n, slices, sizes.p);
    int rc = exclusive_scan_int(ctx, sizes.p, sliceBase.p, slices + 1, nullptr);
    sizes.release();
    if (rc != MOF_OK) return rc;
    int padded = 0;
    MOF_CUDA(read_back(ctx, &padded, sliceBase.p + slices));
    MOF_CUDA(sCol.alloc((size_t)padded));
    MOF_CUDA(sVal.alloc((size_t)padded));
    MOF_LAUNCH(k_csr_to_sell, blocks_for(32ll * slices, B), B, 0, rowptr, col, val, sliceBase.p, n, slices, sCol.p, sVal.p);
    return MOF_OK;
}

int sell_to_csr(mof_ctx* ctx, int n, const int* rowptr, const int* sliceBase, const int* sCol, const double* sVal, int* col, double* val) {
    MOF_LAUNCH(k_sell_to_csr, blocks_for(n, 256), 256, 0, rowptr, sliceBase, sCol, sVal, n, col, val);
    return MOF_OK;
}

}  // namespace mof
