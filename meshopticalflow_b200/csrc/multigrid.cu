// Multilevel preconditioner for the flow system A = s*R D P + w*S (VectorField.h:67), used inside PCG in
// place of the plain Jacobi scaling. The reference factorises A exactly (Eigen SimplicialLDLT); an
// iterative solve needs ~sqrt(cond(A)) Jacobi-PCG iterations (4 500 at 3.1M unknowns), a V-cycle brings
// that down to a few dozen, each costing about five SpMV-sized passes.
//
// Construction (aggregation multigrid on an octree, all on the GPU, deterministic):
//  * The unknowns live on mesh edges; a smooth tangent field looks, locally, like a constant vector of the
//    ambient space, whose Whitney coefficients are c . (x_head - x_tail). So every aggregate carries three
//    coarse unknowns (c_x, c_y, c_z) and the prolongation of an edge is its edge vector.
//  * Aggregates are the occupied cells of a uniform grid over the edge midpoints (level-1 cells hold ~40
//    edges); coarser levels are the parent cells of the octree, down to <= 128 cells, where the dense inverse
//    is applied. Cells are numbered in Morton order, so the children of a cell are contiguous.
//  * Cells are at least as wide as the longest edge, so coupled aggregates are always grid neighbours and every
//    coarse operator is a 27-point stencil of 3x3 blocks: no sparse pattern, no SpGEMM. The Galerkin blocks are
//    re-summed from the current A (they follow the data term) once per flow system; the octree, the
//    neighbour tables and the per-entry stencil slots depend on the mesh only.
//  * V(1,1) cycle, damped Jacobi smoothing (block 3x3 on coarse levels), damping from a Gershgorin bound
//    (guaranteed < 2/rho, which keeps the cycle symmetric positive definite so that plain PCG applies).
// If the mesh does not fit the scheme (an edge longer than a level-1 cell couples non-neighbouring cells) the
// solver falls back to the Jacobi-PCG kernel of pcg_kernels.cu.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "mof_internal.cuh"

namespace mof {

namespace {

constexpr int B = 256;
constexpr int NBLK = kSMs * 4;   // fixed grid of the reduction-producing kernels (deterministic partials)
constexpr int MAXL = 9;          // finest admissible grid level (512^3 cells)
constexpr int SLOT_CENTER = 13;
// Coarse operators are stored component-major: entry k of block (I, slot) at ((slot*9 + k) * N + I), so that one
// thread per cell reads coalesced; block inverses likewise at (k * N + I).
__host__ __device__ __forceinline__ size_t blk(int N, int I, int slot, int k) { return ((size_t)slot * 9 + k) * (size_t)N + I; }

struct MgLevel {
    int gridLevel = 0, N = 0;
    DBuf<int> code, nbr, parent, firstChild;   // firstChild: children (ids on the finer level) of each node, N+1 entries
    DBuf<double> blocks, binv;                 // [N][27][9], [N][9]
    DBuf<double> r, z, t;                      // [3N]
    double omega = 0.6;
};

}  // namespace

struct Multigrid {
    bool usable = false;
    int K = 0;                      // number of coarse levels
    std::vector<MgLevel> lev;       // lev[0] = level 1 (finest aggregates)
    DBuf<double> evec, emid;        // [E][3]
    DBuf<int> agg, aggPtr, aggEdges;
    DBuf<signed char> slotOf;       // per sliced-layout entry of A: stencil slot at level 1, -1 for padding
    DBuf<double> cinv;              // dense inverse on the coarsest level
    DBuf<double> fz, fz2, ft, fr, fp, fq;   // fine-level vectors of the cycle and of PCG
    DBuf<double> partial, scal;
    double omega0 = 0.6;
    int gamma = 2;                  // coarse-grid corrections per level visit
    int gammaLevels = 0;            // ... on the first gammaLevels coarse levels (0 = plain V-cycle, the fastest in time on B200)
    double* hostRR = nullptr;       // pinned: residual norms of the two iterations of one graph replay
    std::vector<double> hostBlocks, hostDense;
    std::vector<int> hostNbr;
};

namespace {

// ---------------------------------------------------------------------------------------- utilities

__host__ __device__ __forceinline__ unsigned spread3(unsigned v) {  // 10 bits -> every third bit
    v &= 0x3ff;
    v = (v | (v << 16)) & 0x030000ff;
    v = (v | (v << 8)) & 0x0300f00f;
    v = (v | (v << 4)) & 0x030c30c3;
    v = (v | (v << 2)) & 0x09249249;
    return v;
}
__host__ __device__ __forceinline__ unsigned compact3(unsigned v) {
    v &= 0x09249249;
    v = (v | (v >> 2)) & 0x030c30c3;
    v = (v | (v >> 4)) & 0x0300f00f;
    v = (v | (v >> 8)) & 0x030000ff;
    v = (v | (v >> 16)) & 0x3ff;
    return v;
}
__host__ __device__ __forceinline__ unsigned morton3(unsigned x, unsigned y, unsigned z) { return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2); }

__global__ void k_edge_geometry(const double* __restrict__ pos, const int* __restrict__ tri, const int* __restrict__ expanded, int E, double* __restrict__ evec,
                                double* __restrict__ emid, double* __restrict__ len2) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int h = expanded[e], t = h / 3, j = h - 3 * t;
    int a = tri[3 * t + (j + 1) % 3], b = tri[3 * t + (j + 2) % 3];
    double l2 = 0;
    for (int k = 0; k < 3; k++) {
        double pa = pos[3 * a + k], pb = pos[3 * b + k];
        evec[3 * e + k] = pb - pa, emid[3 * e + k] = 0.5 * (pa + pb);
        l2 += (pb - pa) * (pb - pa);
    }
    len2[e] = l2;
}

// min (mode 0) / max (mode 1) of a strided array, two stages.
__global__ void k_minmax_partial(const double* __restrict__ in, long long n, int stride, int offset, int mode, double* __restrict__ partial) {
    __shared__ double sh[B];
    double s = mode ? -1e300 : 1e300;
    for (long long i = (long long)blockIdx.x * B + threadIdx.x; i < n; i += (long long)gridDim.x * B) {
        double v = in[i * stride + offset];
        s = mode ? fmax(s, v) : fmin(s, v);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = mode ? fmax(sh[threadIdx.x], sh[threadIdx.x + o]) : fmin(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void k_minmax_final(const double* __restrict__ partial, int np, int mode, double* __restrict__ out) {
    __shared__ double sh[B];
    double s = mode ? -1e300 : 1e300;
    for (int i = threadIdx.x; i < np; i += B) s = mode ? fmax(s, partial[i]) : fmin(s, partial[i]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = mode ? fmax(sh[threadIdx.x], sh[threadIdx.x + o]) : fmin(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

int minmax(mof_ctx* ctx, Multigrid& mg, const double* in, long long n, int stride, int offset, int mode, double* outDevice) {
    MOF_LAUNCH(k_minmax_partial, NBLK, B, 0, in, n, stride, offset, mode, mg.partial.p);
    MOF_LAUNCH(k_minmax_final, 1, B, 0, mg.partial.p, NBLK, mode, outDevice);
    return MOF_OK;
}

// ------------------------------------------------------------------------------- octree construction

struct GridMap {
    double lo[3], inv;  // cell index at level L = floor((p - lo) * inv * 2^L)
};

__device__ __forceinline__ unsigned cell_code(const GridMap& gm, const double* p, int L) {
    unsigned c[3];
    const double s = (double)(1u << L);
    for (int k = 0; k < 3; k++) {
        double v = (p[k] - gm.lo[k]) * gm.inv * s;
        int q = (int)v;
        c[k] = (unsigned)min(max(q, 0), (1 << L) - 1);
    }
    return morton3(c[0], c[1], c[2]);
}

__global__ void k_mark_cells(GridMap gm, const double* __restrict__ emid, int E, int L, int* __restrict__ occ) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    occ[cell_code(gm, emid + 3 * e, L)] = 1;
}
__global__ void k_mark_parents(const int* __restrict__ occ, long long cells, int* __restrict__ occParent) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    if (occ[c]) occParent[c >> 3] = 1;
}
__global__ void k_node_codes(const int* __restrict__ occ, const int* __restrict__ rank, long long cells, int* __restrict__ code) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    if (occ[c]) code[rank[c]] = (int)c;
}
__global__ void k_neighbours(const int* __restrict__ code, const int* __restrict__ occ, const int* __restrict__ rank, int N, int L, int* __restrict__ nbr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 27) return;
    int I = i / 27, s = i - 27 * I;
    unsigned c = (unsigned)code[I];
    int x = (int)compact3(c), y = (int)compact3(c >> 1), z = (int)compact3(c >> 2);
    int nx = x + s / 9 - 1, ny = y + (s / 3) % 3 - 1, nz = z + s % 3 - 1, lim = 1 << L;
    int out = -1;
    if (nx >= 0 && ny >= 0 && nz >= 0 && nx < lim && ny < lim && nz < lim) {
        unsigned m = morton3(nx, ny, nz);
        if (occ[m]) out = rank[m];
    }
    nbr[i] = out;
}
__global__ void k_parents(const int* __restrict__ code, const int* __restrict__ rankParent, int N, int Nparent, int* __restrict__ parent, int* __restrict__ firstChild) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I > N) return;
    if (I == N) { firstChild[Nparent] = N; return; }
    int p = rankParent[(unsigned)code[I] >> 3];
    parent[I] = p;
    if (I == 0 || rankParent[(unsigned)code[I - 1] >> 3] != p) firstChild[p] = I;
}

__global__ void k_edge_aggregate(GridMap gm, const double* __restrict__ emid, const int* __restrict__ rank, int E, int L, int* __restrict__ agg, int* __restrict__ count) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int a = rank[cell_code(gm, emid + 3 * e, L)];
    agg[e] = a;
    atomicAdd(&count[a], 1);
}
__global__ void k_aggregate_fill(const int* __restrict__ agg, int E, int* cursor, int* __restrict__ list) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    list[atomicAdd(&cursor[agg[e]], 1)] = e;
}
// Ascending edge ids within every aggregate (fixed summation order). One warp per aggregate, rank sort.
__global__ void k_aggregate_sort(const int* __restrict__ ptr, int N, const int* __restrict__ in, int* __restrict__ out) {
    int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (I >= N) return;
    int b = ptr[I], n = ptr[I + 1] - b;
    for (int i = lane; i < n; i += 32) {
        int v = in[b + i], r = 0;
        for (int k = 0; k < n; k++) r += in[b + k] < v;
        out[b + r] = v;
    }
}

// Stencil slot of every entry (e,f) of A at level 1; flags[0] is raised when f's cell is not a neighbour of e's.
__global__ void k_entry_slots(const int* __restrict__ wRowptr, const int* __restrict__ sliceBase, const int* __restrict__ wCol, const int* __restrict__ agg,
                              const int* __restrict__ code, int E, int slices, signed char* __restrict__ slotOf, int* __restrict__ flags) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 32 * slices) return;
    const int longest = (sliceBase[(e >> 5) + 1] - sliceBase[e >> 5]) >> 5;
    int len = e < E ? wRowptr[e + 1] - wRowptr[e] : 0;
    unsigned ci = e < E ? (unsigned)code[agg[e]] : 0;
    int x = (int)compact3(ci), y = (int)compact3(ci >> 1), z = (int)compact3(ci >> 2);
    for (int j = 0; j < longest; j++) {
        size_t k = sell_pos(sliceBase, e, j);
        signed char s = -1;
        if (j < len) {
            unsigned cf = (unsigned)code[agg[wCol[k]]];
            int dx = (int)compact3(cf) - x, dy = (int)compact3(cf >> 1) - y, dz = (int)compact3(cf >> 2) - z;
            if (dx < -1 || dx > 1 || dy < -1 || dy > 1 || dz < -1 || dz > 1) flags[0] = 1;
            else s = (signed char)((dx + 1) * 9 + (dy + 1) * 3 + (dz + 1));
        }
        slotOf[k] = s;
    }
}

// ------------------------------------------------------------------------------------ Galerkin values

// Level-1 block (I, slot) = sum over edges e of I, entries f of row e that fall in that neighbour cell, of
// A_ef * v_e v_f^T. One thread per (I, slot); the 27 threads of a cell walk the same entries (broadcast).
__global__ void k_level1_blocks(const int* __restrict__ aggPtr, const int* __restrict__ aggEdges, const int* __restrict__ wRowptr, const int* __restrict__ sliceBase,
                                const int* __restrict__ wCol, const double* __restrict__ wA, const signed char* __restrict__ slotOf, const double* __restrict__ evec,
                                const int* __restrict__ nbr, int N, double* __restrict__ blocks) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 27) return;
    int I = i / 27, s = i - 27 * I;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (nbr[i] >= 0)
        for (int q = aggPtr[I]; q < aggPtr[I + 1]; q++) {
            int e = aggEdges[q], len = wRowptr[e + 1] - wRowptr[e];
            double w[3] = {0, 0, 0};
            for (int j = 0; j < len; j++) {
                size_t k = sell_pos(sliceBase, e, j);
                if (slotOf[k] != s) continue;
                const double a = wA[k];
                const double* vf = evec + 3 * (size_t)wCol[k];
                w[0] += a * vf[0], w[1] += a * vf[1], w[2] += a * vf[2];
            }
            const double* ve = evec + 3 * (size_t)e;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) acc[3 * r + c] += ve[r] * w[c];
        }
    for (int k = 0; k < 9; k++) blocks[blk(N, I, s, k)] = acc[k];
}

// Coarser block (I', slot') = sum of the finer blocks (I, s) with parent(I) = I' and parent(nbr(I, s)) = nbr'(I', slot').
__global__ void k_coarsen_blocks(const int* __restrict__ firstChild, const int* __restrict__ nbrFine, const int* __restrict__ parentFine, const double* __restrict__ blocksFine,
                                 int Nfine, const int* __restrict__ nbrCoarse, int Ncoarse, double* __restrict__ blocksCoarse) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ncoarse * 27) return;
    int Ip = i / 27, sp = i - 27 * Ip;
    int Jp = nbrCoarse[i];
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (Jp >= 0)
        for (int I = firstChild[Ip]; I < firstChild[Ip + 1]; I++)
            for (int s = 0; s < 27; s++) {
                int J = nbrFine[I * 27 + s];
                if (J < 0 || parentFine[J] != Jp) continue;
                for (int k = 0; k < 9; k++) acc[k] += blocksFine[blk(Nfine, I, s, k)];
            }
    for (int k = 0; k < 9; k++) blocksCoarse[blk(Ncoarse, Ip, sp, k)] = acc[k];
}

// Inverse of the (slightly shifted) diagonal block, and the Gershgorin bound of Binv * A for this node.
__global__ void k_block_inverse(const double* __restrict__ blocks, const int* __restrict__ nbr, int N, double* __restrict__ binv, double* __restrict__ bound) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= N) return;
    double m[9];
    for (int k = 0; k < 9; k++) m[k] = blocks[blk(N, I, SLOT_CENTER, k)];
    // Pseudo-inverse through the eigen-decomposition (cyclic Jacobi rotations, accurate for each eigenvalue
    // separately). Cells on the fringe of the surface hold one or two edges, or coplanar ones: their block is
    // rank deficient and the directions with eigenvalue < 1e-8 * largest are simply not smoothed (they prolong to
    // ~0 on the edges anyway). An explicit cofactor inverse of such a block is NOT good enough: its error in
    // the well-conditioned directions scales with the condition number.
    double a00 = m[0], a11 = m[4], a22 = m[8], a01 = 0.5 * (m[1] + m[3]), a02 = 0.5 * (m[2] + m[6]), a12 = 0.5 * (m[5] + m[7]);
    double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // columns = eigenvectors
    for (int sweep = 0; sweep < 8; sweep++) {
#define MOF_ROT(app, aqq, apq, arp, arq, p, q)                                                   \
    if (apq != 0.) {                                                                             \
        double theta = (aqq - app) / (2. * apq);                                                 \
        double t = (theta >= 0 ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1.));           \
        double c = 1. / sqrt(t * t + 1.), s = t * c;                                             \
        double npp = app - t * apq, nqq = aqq + t * apq;                                         \
        double nrp = c * arp - s * arq, nrq = s * arp + c * arq;                                 \
        app = npp, aqq = nqq, apq = 0., arp = nrp, arq = nrq;                                    \
        for (int r = 0; r < 3; r++) {                                                            \
            double vp = V[3 * r + p], vq = V[3 * r + q];                                         \
            V[3 * r + p] = c * vp - s * vq, V[3 * r + q] = s * vp + c * vq;                      \
        }                                                                                        \
    }
        MOF_ROT(a00, a11, a01, a02, a12, 0, 1)
        MOF_ROT(a00, a22, a02, a01, a12, 0, 2)
        MOF_ROT(a11, a22, a12, a01, a02, 1, 2)
#undef MOF_ROT
    }
    double lam[3] = {a00, a11, a22};
    double top = fmax(lam[0], fmax(lam[1], lam[2]));
    double inv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int e = 0; e < 3; e++) {
        if (!(lam[e] > 1e-8 * top) || !(top > 0)) continue;
        double il = 1. / lam[e];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) inv[3 * r + c] += il * V[3 * r + e] * V[3 * c + e];
    }
    for (int k = 0; k < 9; k++) binv[(size_t)k * N + I] = inv[k];
    double rows[3] = {0, 0, 0};
    for (int s = 0; s < 27; s++) {
        if (nbr[I * 27 + s] < 0) continue;
        double b[9];
        for (int k = 0; k < 9; k++) b[k] = blocks[blk(N, I, s, k)];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) rows[r] += fabs(inv[3 * r] * b[c] + inv[3 * r + 1] * b[3 + c] + inv[3 * r + 2] * b[6 + c]);
    }
    bound[I] = fmax(rows[0], fmax(rows[1], rows[2]));
}

// Gershgorin bound of D^-1 A on the fine level: max over rows of sum_j |a_ej| / a_ee.
__global__ void k_fine_bound(const int* __restrict__ wRowptr, const int* __restrict__ sliceBase, const double* __restrict__ wA, const double* __restrict__ dinv, int E,
                             double* __restrict__ bound) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int len = wRowptr[e + 1] - wRowptr[e];
    double s = 0;
    for (int j = 0; j < len; j++) s += fabs(wA[sell_pos(sliceBase, e, j)]);
    bound[e] = s * fabs(dinv[e]);
}

// ------------------------------------------------------------------------------------- cycle kernels

__global__ void k_fine_presmooth(const double* __restrict__ r, const double* __restrict__ dinv, double omega, int E, double* __restrict__ z) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) z[e] = omega * dinv[e] * r[e];
}

// Sliced SpMV (warp = slice, lane = row), as in pcg_kernels.cu. mode 0: out = A in ; mode 1: out = b - A in ;
// mode 2: out = in + omega * dinv * (b - A in)  (one damped Jacobi sweep).
constexpr int BATCH = 6;
__global__ void __launch_bounds__(B) k_fine_apply(int n, const int* __restrict__ sliceBase, const int* __restrict__ col, const double* __restrict__ val,
                                                 const double* __restrict__ b, const double* __restrict__ dinv, double omega, const double* __restrict__ in,
                                                 double* __restrict__ out, int mode) {
    const int lane = threadIdx.x & 31;
    const int slices = (n + 31) >> 5;
    const int warps = gridDim.x * (B / 32);
    for (int s = blockIdx.x * (B / 32) + (threadIdx.x >> 5); s < slices; s += warps) {
        const int base = sliceBase[s];
        const int len = (sliceBase[s + 1] - base) >> 5;
        const double* v0 = val + (size_t)base + lane;
        const int* c0 = col + (size_t)base + lane;
        const int row = 32 * s + lane;
        double acc = 0;
        for (int j0 = 0; j0 < len; j0 += BATCH) {
            double v[BATCH], x[BATCH];
            int c[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; u++) {
                bool ok = j0 + u < len;
                v[u] = ok ? __ldcs(v0 + 32 * (size_t)(j0 + u)) : 0.;
                c[u] = ok ? __ldcs(c0 + 32 * (size_t)(j0 + u)) : 0;
            }
#pragma unroll
            for (int u = 0; u < BATCH; u++) x[u] = in[c[u]];
#pragma unroll
            for (int u = 0; u < BATCH; u++)
                if (j0 + u < len) acc += v[u] * x[u];
        }
        if (row < n) {
            if (mode == 0) out[row] = acc;
            else if (mode == 1) out[row] = b[row] - acc;
            else out[row] = in[row] + omega * dinv[row] * (b[row] - acc);
        }
    }
}

// rc[I] = sum over the edges of aggregate I of v_e * r_e   (P1^T r)
__global__ void k_restrict_fine(const int* __restrict__ aggPtr, const int* __restrict__ aggEdges, const double* __restrict__ evec, const double* __restrict__ r, int N,
                                double* __restrict__ rc) {
    // one warp per aggregate: lanes stride over its edges, then a fixed shuffle tree
    int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (I >= N) return;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int q = aggPtr[I] + lane; q < aggPtr[I + 1]; q += 32) {
        int e = aggEdges[q];
        double re = r[e];
        a0 += evec[3 * (size_t)e] * re, a1 += evec[3 * (size_t)e + 1] * re, a2 += evec[3 * (size_t)e + 2] * re;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o), a1 += __shfl_xor_sync(0xffffffffu, a1, o), a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) rc[3 * (size_t)I] = a0, rc[3 * (size_t)I + 1] = a1, rc[3 * (size_t)I + 2] = a2;
}
// z_e += v_e . zc[agg(e)]   (P1 zc)
__global__ void k_prolong_fine(const int* __restrict__ agg, const double* __restrict__ evec, const double* __restrict__ zc, int E, double* __restrict__ z) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const double* c = zc + 3 * (size_t)agg[e];
    z[e] += evec[3 * (size_t)e] * c[0] + evec[3 * (size_t)e + 1] * c[1] + evec[3 * (size_t)e + 2] * c[2];
}

__device__ __forceinline__ void mat3_vec(const double* m, const double* v, double* out) {
    out[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    out[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    out[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
__global__ void k_coarse_presmooth(const double* __restrict__ binv, const double* __restrict__ r, double omega, int N, double* __restrict__ z) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= N) return;
    double m[9], o[3];
    for (int k = 0; k < 9; k++) m[k] = binv[(size_t)k * N + I];
    mat3_vec(m, r + 3 * (size_t)I, o);
    for (int k = 0; k < 3; k++) z[3 * (size_t)I + k] = omega * o[k];
}
// mode 1: out = r - A z ; mode 2: out = z + omega * Binv (r - A z). One thread per node, 27 block products.
__global__ void k_coarse_apply(const double* __restrict__ blocks, const int* __restrict__ nbr, const double* __restrict__ binv, const double* __restrict__ r,
                               const double* __restrict__ z, double omega, int N, int mode, double* __restrict__ out) {
    // CTA = 27 warps: warp = stencil slot, lane = cell (32 consecutive cells), so every block component is read as
    // 32 consecutive words and all 27 slots of a cell are in flight at once; the 27 partial products are then summed
    // in slot order by the first warp.
    __shared__ double part[27][32][3];
    const int slot = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int I = blockIdx.x * 32 + lane;
    double o[3] = {0, 0, 0};
    if (I < N) {
        int J = nbr[I * 27 + slot];
        if (J >= 0) {
            double m[9];
#pragma unroll
            for (int k = 0; k < 9; k++) m[k] = blocks[blk(N, I, slot, k)];
            mat3_vec(m, z + 3 * (size_t)J, o);
        }
    }
    part[slot][lane][0] = o[0], part[slot][lane][1] = o[1], part[slot][lane][2] = o[2];
    __syncthreads();
    if (slot != 0 || I >= N) return;
    double acc[3] = {0, 0, 0};
    for (int s = 0; s < 27; s++) acc[0] += part[s][lane][0], acc[1] += part[s][lane][1], acc[2] += part[s][lane][2];
    double res[3] = {r[3 * (size_t)I] - acc[0], r[3 * (size_t)I + 1] - acc[1], r[3 * (size_t)I + 2] - acc[2]};
    if (mode == 1) {
        for (int k = 0; k < 3; k++) out[3 * (size_t)I + k] = res[k];
    } else {
        double m[9], o[3];
        for (int k = 0; k < 9; k++) m[k] = binv[(size_t)k * N + I];
        mat3_vec(m, res, o);
        for (int k = 0; k < 3; k++) out[3 * (size_t)I + k] = z[3 * (size_t)I + k] + omega * o[k];
    }
}
__global__ void k_restrict_coarse(const int* __restrict__ firstChild, const double* __restrict__ rFine, int Ncoarse, double* __restrict__ rc) {
    int Ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (Ip >= Ncoarse) return;
    double a[3] = {0, 0, 0};
    for (int I = firstChild[Ip]; I < firstChild[Ip + 1]; I++)
        for (int k = 0; k < 3; k++) a[k] += rFine[3 * (size_t)I + k];
    for (int k = 0; k < 3; k++) rc[3 * (size_t)Ip + k] = a[k];
}
__global__ void k_prolong_coarse(const int* __restrict__ parent, const double* __restrict__ zc, int N, double* __restrict__ z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * N) return;
    z[i] += zc[3 * (size_t)parent[i / 3] + i % 3];
}
__global__ void k_dense_apply(const double* __restrict__ m, const double* __restrict__ r, int n, double* __restrict__ z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0;
    for (int k = 0; k < n; k++) s += m[(size_t)i * n + k] * r[k];
    z[i] = s;
}

// ------------------------------------------------------------------------------------- PCG kernels

enum { S_RZ = 0, S_PQ = 1, S_ALPHA = 2, S_BETA = 3, S_RR = 4, S_BB = 5, S_RZNEW = 6 };

__global__ void k_dot_partial(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ partial) {
    __shared__ double sh[B];
    double s = 0;
    for (int i = blockIdx.x * B + threadIdx.x; i < n; i += gridDim.x * B) s += a[i] * b[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// Folds the partials (fixed order) into scal[slot] and derives the PCG scalar that depends on it.
__global__ void k_fold(const double* __restrict__ partial, int np, int slot, double* __restrict__ scal) {
    __shared__ double sh[B];
    double s = 0;
    for (int i = threadIdx.x; i < np; i += B) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double v = sh[0];
        scal[slot] = v;
        if (slot == S_PQ) scal[S_ALPHA] = v != 0 ? scal[S_RZ] / v : 0.;
        if (slot == S_RZNEW) {
            scal[S_BETA] = scal[S_RZ] != 0 ? v / scal[S_RZ] : 0.;
            scal[S_RZ] = v;
        }
    }
}
// x += alpha p ; r -= alpha q ; partial(r.r)
__global__ void k_update_xr(const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ scal, int n, double* __restrict__ x,
                            double* __restrict__ r, double* __restrict__ partial) {
    __shared__ double sh[B];
    const double alpha = scal[S_ALPHA];
    double s = 0;
    for (int i = blockIdx.x * B + threadIdx.x; i < n; i += gridDim.x * B) {
        double rv = r[i] - alpha * q[i];
        x[i] += alpha * p[i];
        r[i] = rv;
        s += rv * rv;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// Power iteration support (spectral radius of Minv A per level, for the Jacobi damping).
__global__ void k_pseudo_random(int n, double* __restrict__ v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = (unsigned)i * 2654435761u + 12345u;
    h ^= h >> 15, h *= 2246822519u, h ^= h >> 13;
    v[i] = (double)(h & 0xffff) / 32768. - 1.;
}
// v = t / sqrt(scal[slot])
__global__ void k_normalise(const double* __restrict__ t, const double* __restrict__ scal, int slot, int n, double* __restrict__ v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = scal[slot];
    v[i] = s > 0 ? t[i] / sqrt(s) : 0.;
}

__global__ void k_direction(const double* __restrict__ z, const double* __restrict__ scal, int n, int first, double* __restrict__ p) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p[i] = first ? z[i] : z[i] + scal[S_BETA] * p[i];
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side

void mg_destroy(mof_ctx* ctx) {
    if (!ctx->mg) return;
    Multigrid& mg = *ctx->mg;
    for (MgLevel& l : mg.lev) {
        l.code.release(), l.nbr.release(), l.parent.release(), l.firstChild.release(), l.blocks.release(), l.binv.release(), l.r.release(), l.z.release(), l.t.release();
    }
    mg.evec.release(), mg.emid.release(), mg.agg.release(), mg.aggPtr.release(), mg.aggEdges.release(), mg.slotOf.release(), mg.cinv.release();
    mg.fz.release(), mg.fz2.release(), mg.ft.release(), mg.fr.release(), mg.fp.release(), mg.fq.release(), mg.partial.release(), mg.scal.release();
    if (mg.hostRR) cudaFreeHost(mg.hostRR);
    delete ctx->mg;
    ctx->mg = nullptr;
}

static int env_int(const char* name, int fallback) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : fallback;
}

// Mesh-dependent part: octree over the edge midpoints, aggregates, neighbour tables, stencil slots.
int mg_setup_mesh(mof_ctx* ctx) {
    mg_destroy(ctx);
    ctx->mg = new Multigrid();
    Multigrid& mg = *ctx->mg;
    if (env_int("MOF_FLOW_MG", 1) == 0) return MOF_OK;
    mg.gamma = std::max(1, std::min(2, env_int("MOF_MG_GAMMA", 2)));
    mg.gammaLevels = std::max(0, env_int("MOF_MG_GAMMA_LEVELS", 0));
    MOF_CUDA(cudaHostAlloc((void**)&mg.hostRR, 8 * sizeof(double), cudaHostAllocDefault));
    const int E = ctx->E;
    MOF_CUDA(mg.evec.alloc(3ull * E));
    MOF_CUDA(mg.emid.alloc(3ull * E));
    MOF_CUDA(mg.partial.alloc(8192));  // per-CTA partials: NBLK of ours, or the persistent-grid size of k_spmv_dot
    MOF_CUDA(mg.scal.alloc(64));
    MOF_CUDA(ctx->dtmp0.reserve((size_t)E));
    MOF_LAUNCH(k_edge_geometry, blocks_for(E, B), B, 0, ctx->pos.p, ctx->tri.p, ctx->expanded.p, E, mg.evec.p, mg.emid.p, ctx->dtmp0.p);
    for (int k = 0; k < 3; k++) {
        MOF_TRY(minmax(ctx, mg, mg.emid.p, E, 3, k, 0, mg.scal.p + k));
        MOF_TRY(minmax(ctx, mg, mg.emid.p, E, 3, k, 1, mg.scal.p + 3 + k));
    }
    MOF_TRY(minmax(ctx, mg, ctx->dtmp0.p, E, 1, 0, 1, mg.scal.p + 6));
    double h[7];
    MOF_CUDA(cudaMemcpyAsync(h, mg.scal.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    double ext = 0;
    for (int k = 0; k < 3; k++) ext = std::max(ext, h[3 + k] - h[k]);
    const double maxEdge = std::sqrt(h[6]);
    if (!(ext > 0) || !(maxEdge > 0)) return MOF_OK;
    ext *= 1.0001;
    GridMap gm;
    for (int k = 0; k < 3; k++) gm.lo[k] = h[k];
    gm.inv = 1. / ext;
    // finest admissible level: cells at least as wide as the longest edge
    int Lmax = 1;
    while (Lmax < MAXL && ext / (double)(1 << (Lmax + 1)) >= maxEdge) Lmax++;
    // occupancy of a surface grows ~4x per level: probe one level, extrapolate to ~target edges per cell
    const int target = env_int("MOF_MG_TARGET", 40);
    const int probe = std::min(Lmax, 5);
    DBuf<int> occ[MAXL + 1], rank[MAXL + 1];
    auto freeAll = [&]() { for (int L = 0; L <= MAXL; L++) occ[L].release(), rank[L].release(); };
    {
        long long cells = 1ll << (3 * probe);
        MOF_CUDA(occ[probe].alloc(cells + 1));
        MOF_CUDA(cudaMemsetAsync(occ[probe].p, 0, sizeof(int) * (cells + 1), ctx->stream));
        MOF_LAUNCH(k_mark_cells, blocks_for(E, B), B, 0, gm, mg.emid.p, E, probe, occ[probe].p);
        MOF_CUDA(rank[probe].alloc(cells + 1));
        int rc = exclusive_scan_int(ctx, occ[probe].p, rank[probe].p, (int)cells + 1, nullptr);
        if (rc != MOF_OK) { freeAll(); return rc; }
    }
    int nProbe = 0;
    MOF_CUDA(cudaMemcpy(&nProbe, rank[probe].p + (1ll << (3 * probe)), sizeof(int), cudaMemcpyDeviceToHost));
    occ[probe].release(), rank[probe].release();
    int L1 = probe + (int)std::lround(std::log((double)E / ((double)target * std::max(nProbe, 1))) / std::log(4.0));
    L1 = std::max(2, std::min(Lmax, L1));
    if (Lmax < 2) return MOF_OK;

    // occupancy and Morton ranks of level L1 and of every coarser level
    for (int L = L1; L >= 1; L--) {
        long long cells = 1ll << (3 * L);
        MOF_CUDA(occ[L].alloc(cells + 1));
        MOF_CUDA(rank[L].alloc(cells + 1));
        MOF_CUDA(cudaMemsetAsync(occ[L].p, 0, sizeof(int) * (cells + 1), ctx->stream));
        if (L == L1) MOF_LAUNCH(k_mark_cells, blocks_for(E, B), B, 0, gm, mg.emid.p, E, L, occ[L].p);
        else MOF_LAUNCH(k_mark_parents, blocks_for(1ll << (3 * (L + 1)), B), B, 0, occ[L + 1].p, 1ll << (3 * (L + 1)), occ[L].p);
        int rc = exclusive_scan_int(ctx, occ[L].p, rank[L].p, (int)cells + 1, nullptr);
        if (rc != MOF_OK) { freeAll(); return rc; }
    }
    std::vector<int> counts(L1 + 1, 0);
    for (int L = L1; L >= 1; L--) MOF_CUDA(cudaMemcpy(&counts[L], rank[L].p + (1ll << (3 * L)), sizeof(int), cudaMemcpyDeviceToHost));
    // coarsest level: the first one with at most 48 cells (144 unknowns, inverted densely on the host)
    int Lc = L1;
    while (Lc > 1 && counts[Lc] > 48) Lc--;
    if (counts[Lc] > 128) { freeAll(); return MOF_OK; }  // pathological: give up, Jacobi-PCG remains
    mg.K = L1 - Lc + 1;
    mg.lev.resize(mg.K);
    for (int l = 0; l < mg.K; l++) {
        MgLevel& lv = mg.lev[l];
        int L = L1 - l;
        lv.gridLevel = L, lv.N = counts[L];
        long long cells = 1ll << (3 * L);
        MOF_CUDA(lv.code.alloc(lv.N));
        MOF_CUDA(lv.nbr.alloc(27ull * lv.N));
        MOF_CUDA(lv.blocks.alloc(27ull * 9 * lv.N));
        MOF_CUDA(lv.binv.alloc(9ull * lv.N));
        MOF_CUDA(lv.r.alloc(3ull * lv.N));
        MOF_CUDA(lv.z.alloc(3ull * lv.N));
        MOF_CUDA(lv.t.alloc(3ull * lv.N));
        MOF_LAUNCH(k_node_codes, blocks_for(cells, B), B, 0, occ[L].p, rank[L].p, cells, lv.code.p);
        MOF_LAUNCH(k_neighbours, blocks_for(27ll * lv.N, B), B, 0, lv.code.p, occ[L].p, rank[L].p, lv.N, L, lv.nbr.p);
    }
    for (int l = 0; l + 1 < mg.K; l++) {
        MgLevel& lv = mg.lev[l];
        MgLevel& up = mg.lev[l + 1];
        MOF_CUDA(lv.parent.alloc(lv.N));
        MOF_CUDA(up.firstChild.alloc(up.N + 1));
        MOF_LAUNCH(k_parents, blocks_for(lv.N + 1, B), B, 0, lv.code.p, rank[up.gridLevel].p, lv.N, up.N, lv.parent.p, up.firstChild.p);
    }
    // level-1 aggregates of the edges
    MgLevel& l1 = mg.lev[0];
    MOF_CUDA(mg.agg.alloc(E));
    MOF_CUDA(mg.aggPtr.alloc(l1.N + 1));
    MOF_CUDA(mg.aggEdges.alloc(E));
    DBuf<int> cnt, cursor, unsorted;
    MOF_CUDA(cnt.alloc(l1.N + 1));
    MOF_CUDA(cursor.alloc(l1.N + 1));
    MOF_CUDA(unsorted.alloc(E));
    MOF_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int) * (l1.N + 1), ctx->stream));
    MOF_LAUNCH(k_edge_aggregate, blocks_for(E, B), B, 0, gm, mg.emid.p, rank[L1].p, E, L1, mg.agg.p, cnt.p);
    int rc = exclusive_scan_int(ctx, cnt.p, mg.aggPtr.p, l1.N + 1, nullptr);
    if (rc == MOF_OK) {
        cudaMemcpyAsync(cursor.p, mg.aggPtr.p, sizeof(int) * (l1.N + 1), cudaMemcpyDeviceToDevice, ctx->stream);
        k_aggregate_fill<<<blocks_for(E, B), B, 0, ctx->stream>>>(mg.agg.p, E, cursor.p, unsorted.p);
        k_aggregate_sort<<<blocks_for(32ll * l1.N, B), B, 0, ctx->stream>>>(mg.aggPtr.p, l1.N, unsorted.p, mg.aggEdges.p);
        ctx->stats.kernelLaunches += 2;
    }
    // stencil slot of every entry of A
    MOF_CUDA(mg.slotOf.alloc((size_t)ctx->wPadded));
    MOF_CUDA(cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * 16, ctx->stream));
    MOF_LAUNCH(k_entry_slots, blocks_for(32ll * ctx->wSlices, B), B, 0, ctx->wRowptr.p, ctx->wSliceBase.p, ctx->wCol.p, mg.agg.p, l1.code.p, E, ctx->wSlices, mg.slotOf.p,
               ctx->flags.p);
    int hflag = 0;
    MOF_CUDA(cudaMemcpyAsync(&hflag, ctx->flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    cnt.release(), cursor.release(), unsorted.release();
    freeAll();
    if (rc != MOF_OK) return rc;
    if (hflag) return MOF_OK;  // coupled aggregates that are not neighbours: stay with Jacobi
    MOF_CUDA(mg.fz.alloc(E));
    MOF_CUDA(mg.fz2.alloc(E));
    MOF_CUDA(mg.ft.alloc(E));
    MOF_CUDA(mg.fr.alloc(E));
    MOF_CUDA(mg.fp.alloc(E));
    MOF_CUDA(mg.fq.alloc(E));
    const int nc = 3 * mg.lev.back().N;
    MOF_CUDA(mg.cinv.alloc((size_t)nc * nc));
    mg.usable = true;
    return MOF_OK;
}

bool mg_usable(const mof_ctx* ctx) { return ctx->mg && ctx->mg->usable; }

// Value-dependent part, once per flow system: Galerkin blocks on every level, block inverses, damping factors,
// dense inverse on the coarsest level.
int mg_update_values(mof_ctx* ctx) {
    Multigrid& mg = *ctx->mg;
    const int E = ctx->E;
    MgLevel& l1 = mg.lev[0];
    MOF_LAUNCH(k_level1_blocks, blocks_for(27ll * l1.N, B), B, 0, mg.aggPtr.p, mg.aggEdges.p, ctx->wRowptr.p, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, mg.slotOf.p,
               mg.evec.p, l1.nbr.p, l1.N, l1.blocks.p);
    for (int l = 0; l + 1 < mg.K; l++) {
        MgLevel& lv = mg.lev[l];
        MgLevel& up = mg.lev[l + 1];
        MOF_LAUNCH(k_coarsen_blocks, blocks_for(27ll * up.N, B), B, 0, up.firstChild.p, lv.nbr.p, lv.parent.p, lv.blocks.p, lv.N, up.nbr.p, up.N, up.blocks.p);
    }
    // Damping of the Jacobi smoothers: omega_l = 1.4 / rho_l with rho_l the spectral radius of Minv A on that level,
    // estimated by power iteration on I + Minv A (all eigenvalues of Minv A are positive). The fine level is
    // additionally capped by its Gershgorin bound, which is guaranteed. omega * rho < 2 keeps the cycle positive
    // definite; should an estimate ever be too low, PCG fails to converge and update_flow falls back to Jacobi-PCG.
    const int powerIts = 10;
    MOF_CUDA(ctx->dtmp0.reserve((size_t)E));
    MOF_LAUNCH(k_fine_bound, blocks_for(E, B), B, 0, ctx->wRowptr.p, ctx->wSliceBase.p, ctx->wA.p, ctx->wDinv.p, E, ctx->dtmp0.p);
    MOF_TRY(minmax(ctx, mg, ctx->dtmp0.p, E, 1, 0, 1, mg.scal.p + 8));
    {
        MOF_CUDA(cudaMemsetAsync(mg.ft.p, 0, sizeof(double) * E, ctx->stream));
        MOF_LAUNCH(k_pseudo_random, blocks_for(E, B), B, 0, E, mg.fz.p);
        for (int it = 0; it <= powerIts; it++) {
            MOF_LAUNCH(k_dot_partial, NBLK, B, 0, mg.fz.p, mg.fz.p, E, mg.partial.p);
            MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, 16, mg.scal.p);
            if (it == powerIts) break;
            MOF_LAUNCH(k_normalise, blocks_for(E, B), B, 0, mg.fz.p, mg.scal.p, 16, E, mg.fz.p);
            MOF_LAUNCH(k_fine_apply, kSMs * 8, B, 0, E, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, mg.ft.p, ctx->wDinv.p, -1., mg.fz.p, mg.fz2.p, 2);
            std::swap(mg.fz.p, mg.fz2.p);
        }
    }
    for (int l = 0; l < mg.K; l++) {
        MgLevel& lv = mg.lev[l];
        MOF_CUDA(ctx->dtmp0.reserve((size_t)lv.N));
        MOF_LAUNCH(k_block_inverse, blocks_for(lv.N, B), B, 0, lv.blocks.p, lv.nbr.p, lv.N, lv.binv.p, ctx->dtmp0.p);
        if (l == mg.K - 1) break;
        const int n3 = 3 * lv.N;
        MOF_CUDA(cudaMemsetAsync(lv.r.p, 0, sizeof(double) * n3, ctx->stream));
        MOF_LAUNCH(k_pseudo_random, blocks_for(n3, B), B, 0, n3, lv.z.p);
        for (int it = 0; it <= powerIts; it++) {
            MOF_LAUNCH(k_dot_partial, NBLK, B, 0, lv.z.p, lv.z.p, n3, mg.partial.p);
            MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, 17 + l, mg.scal.p);
            if (it == powerIts) break;
            MOF_LAUNCH(k_normalise, blocks_for(n3, B), B, 0, lv.z.p, mg.scal.p, 17 + l, n3, lv.z.p);
            MOF_LAUNCH(k_coarse_apply, blocks_for(lv.N, 32), 27 * 32, 0, lv.blocks.p, lv.nbr.p, lv.binv.p, lv.r.p, lv.z.p, -1., lv.N, 2, lv.t.p);
            std::swap(lv.z.p, lv.t.p);
        }
    }
    // coarsest level: dense matrix on the host, Cholesky inverse, back to the device
    MgLevel& lc = mg.lev.back();
    const int Nc = lc.N, nc = 3 * Nc;
    mg.hostBlocks.resize((size_t)Nc * 27 * 9);
    mg.hostNbr.resize((size_t)Nc * 27);
    double hs[40] = {0};
    MOF_CUDA(cudaMemcpyAsync(mg.hostBlocks.data(), lc.blocks.p, sizeof(double) * mg.hostBlocks.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(mg.hostNbr.data(), lc.nbr.p, sizeof(int) * mg.hostNbr.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(hs, mg.scal.p, sizeof(double) * 40, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    const char* fixedOmega = getenv("MOF_MG_OMEGA");
    auto damping = [&](double normSq) {
        double rho = std::sqrt(std::max(normSq, 0.)) - 1.;  // the iterate was normalised before the last application of I + Minv A
        if (fixedOmega) return atof(fixedOmega);
        return rho > 0.1 ? std::min(0.8, 1.4 / rho) : 0.6;
    };
    mg.omega0 = damping(hs[16]);
    if (!fixedOmega && hs[8] > 0) mg.omega0 = std::min(mg.omega0, 1.9 / hs[8]);
    for (int l = 0; l + 1 < mg.K; l++) mg.lev[l].omega = damping(hs[17 + l]);
    if (env_int("MOF_MG_VERBOSE", 0)) {
        fprintf(stderr, "[mg] fine: E=%d rho %.3f gershgorin %.3f omega %.3f\n", E, std::sqrt(hs[16]) - 1., hs[8], mg.omega0);
        for (int l = 0; l < mg.K; l++)
            fprintf(stderr, "[mg] level %d: grid 2^%d, %d cells, rho %.3f omega %.3f\n", l + 1, mg.lev[l].gridLevel, mg.lev[l].N, l + 1 < mg.K ? std::sqrt(hs[17 + l]) - 1. : 0.,
                    mg.lev[l].omega);
    }
    if (env_int("MOF_MG_VERBOSE", 0) >= 2) {  // consistency of the level-1 operator: symmetry, positive diagonal blocks
        MgLevel& l1 = mg.lev[0];
        std::vector<double> hb((size_t)l1.N * 243);
        std::vector<int> hn((size_t)l1.N * 27);
        cudaMemcpy(hb.data(), l1.blocks.p, sizeof(double) * hb.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(hn.data(), l1.nbr.p, sizeof(int) * hn.size(), cudaMemcpyDeviceToHost);
        double asym = 0, scale = 0, minRatio = 1e300;
        int badDiag = 0;
        for (int I = 0; I < l1.N; I++) {
            for (int s = 0; s < 27; s++) {
                int J = hn[(size_t)I * 27 + s];
                if (J < 0) continue;
                for (int k = 0; k < 9; k++) {
                    double a = hb[blk(l1.N, I, s, k)], bT = hb[blk(l1.N, J, 26 - s, (k % 3) * 3 + k / 3)];
                    asym = std::max(asym, std::fabs(a - bT)), scale = std::max(scale, std::fabs(a));
                }
            }
            double m[9];
            for (int k = 0; k < 9; k++) m[k] = hb[blk(l1.N, I, SLOT_CENTER, k)];
            double d1 = m[0], d2 = m[0] * m[4] - m[1] * m[3];
            double d3 = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
            if (!(d1 > 0 && d2 > 0 && d3 > 0)) badDiag++;
            double tr = m[0] + m[4] + m[8];
            minRatio = std::min(minRatio, d3 / (tr * tr * tr / 27.));
        }
        fprintf(stderr, "[mg] level-1 check: max |B(I,J) - B(J,I)^T| = %.3e (scale %.3e), non-positive diagonal blocks %d / %d, min det/(tr/3)^3 = %.3e\n", asym, scale, badDiag,
                l1.N, minRatio);
    }
    std::vector<double>& M = mg.hostDense;
    M.assign((size_t)nc * nc, 0.);
    for (int I = 0; I < Nc; I++)
        for (int s = 0; s < 27; s++) {
            int J = mg.hostNbr[(size_t)I * 27 + s];
            if (J < 0) continue;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) M[(size_t)(3 * I + r) * nc + 3 * J + c] += mg.hostBlocks[blk(Nc, I, s, 3 * r + c)];
        }
    double trace = 0;
    for (int i = 0; i < nc; i++) trace += M[(size_t)i * nc + i];
    for (int i = 0; i < nc; i++) {
        M[(size_t)i * nc + i] += 1e-12 * trace / nc;
        for (int j = 0; j < i; j++) M[(size_t)i * nc + j] = M[(size_t)j * nc + i] = 0.5 * (M[(size_t)i * nc + j] + M[(size_t)j * nc + i]);
    }
    // Cholesky M = L L^T (in place, lower), then inverse = L^-T L^-1
    std::vector<double> Lm(M);
    for (int j = 0; j < nc; j++) {
        double d = Lm[(size_t)j * nc + j];
        for (int k = 0; k < j; k++) d -= Lm[(size_t)j * nc + k] * Lm[(size_t)j * nc + k];
        if (!(d > 0)) { mg.usable = false; return MOF_OK; }  // not positive definite: Jacobi-PCG takes over
        d = std::sqrt(d);
        Lm[(size_t)j * nc + j] = d;
        for (int i = j + 1; i < nc; i++) {
            double s = Lm[(size_t)i * nc + j];
            for (int k = 0; k < j; k++) s -= Lm[(size_t)i * nc + k] * Lm[(size_t)j * nc + k];
            Lm[(size_t)i * nc + j] = s / d;
        }
    }
    std::vector<double> Li((size_t)nc * nc, 0.);  // L^-1, lower triangular
    for (int c = 0; c < nc; c++) {
        Li[(size_t)c * nc + c] = 1. / Lm[(size_t)c * nc + c];
        for (int i = c + 1; i < nc; i++) {
            double s = 0;
            for (int k = c; k < i; k++) s -= Lm[(size_t)i * nc + k] * Li[(size_t)k * nc + c];
            Li[(size_t)i * nc + c] = s / Lm[(size_t)i * nc + i];
        }
    }
    for (int i = 0; i < nc; i++)
        for (int j = 0; j <= i; j++) {
            double s = 0;
            for (int k = i; k < nc; k++) s += Li[(size_t)k * nc + i] * Li[(size_t)k * nc + j];
            M[(size_t)i * nc + j] = M[(size_t)j * nc + i] = s;
        }
    MOF_CUDA(cudaMemcpyAsync(mg.cinv.p, M.data(), sizeof(double) * (size_t)nc * nc, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

// One V(1,1) cycle on the coarse hierarchy: lev[l].r in, lev[l].z out.
static int coarse_cycle(mof_ctx* ctx, Multigrid& mg, int l) {
    MgLevel& lv = mg.lev[l];
    if (l == mg.K - 1) {
        MOF_LAUNCH(k_dense_apply, blocks_for(3 * lv.N, 128), 128, 0, mg.cinv.p, lv.r.p, 3 * lv.N, lv.z.p);
        return MOF_OK;
    }
    MgLevel& up = mg.lev[l + 1];
    MOF_LAUNCH(k_coarse_presmooth, blocks_for(lv.N, B), B, 0, lv.binv.p, lv.r.p, lv.omega, lv.N, lv.z.p);
    // gamma coarse-grid corrections per visit (1 = V-cycle, 2 = W-cycle; the coarse levels are cheap, so W by default)
    for (int g = 0; g < (l < mg.gammaLevels ? mg.gamma : 1); g++) {
        MOF_LAUNCH(k_coarse_apply, blocks_for(lv.N, 32), 27 * 32, 0, lv.blocks.p, lv.nbr.p, lv.binv.p, lv.r.p, lv.z.p, lv.omega, lv.N, 1, lv.t.p);
        MOF_LAUNCH(k_restrict_coarse, blocks_for(up.N, B), B, 0, up.firstChild.p, lv.t.p, up.N, up.r.p);
        MOF_TRY(coarse_cycle(ctx, mg, l + 1));
        MOF_LAUNCH(k_prolong_coarse, blocks_for(3ll * lv.N, B), B, 0, lv.parent.p, up.z.p, lv.N, lv.z.p);
    }
    MOF_LAUNCH(k_coarse_apply, blocks_for(lv.N, 32), 27 * 32, 0, lv.blocks.p, lv.nbr.p, lv.binv.p, lv.r.p, lv.z.p, lv.omega, lv.N, 2, lv.t.p);
    std::swap(lv.z.p, lv.t.p);
    return MOF_OK;
}

// z = V-cycle(r) on the fine level; result in mg.fz.
static int fine_cycle(mof_ctx* ctx, Multigrid& mg, const double* r) {
    const int E = ctx->E;
    const int grid = kSMs * 8;
    MgLevel& l1 = mg.lev[0];
    MOF_LAUNCH(k_fine_presmooth, blocks_for(E, B), B, 0, r, ctx->wDinv.p, mg.omega0, E, mg.fz.p);
    MOF_LAUNCH(k_fine_apply, grid, B, 0, E, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, r, ctx->wDinv.p, mg.omega0, mg.fz.p, mg.ft.p, 1);
    MOF_LAUNCH(k_restrict_fine, blocks_for(32ll * l1.N, B), B, 0, mg.aggPtr.p, mg.aggEdges.p, mg.evec.p, mg.ft.p, l1.N, l1.r.p);
    MOF_TRY(coarse_cycle(ctx, mg, 0));
    MOF_LAUNCH(k_prolong_fine, blocks_for(E, B), B, 0, mg.agg.p, mg.evec.p, l1.z.p, E, mg.fz.p);
    MOF_LAUNCH(k_fine_apply, grid, B, 0, E, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, r, ctx->wDinv.p, mg.omega0, mg.fz.p, mg.fz2.p, 2);
    std::swap(mg.fz.p, mg.fz2.p);
    return MOF_OK;
}

// k_spmv_dot of pcg_kernels.cu: q = A p with per-CTA partials of p.q
int spmv_dot_launch(mof_ctx* ctx, int n, const int* sliceBase, const int* col, const double* val, const double* x, double* y, double* partial, int* gridOut);

// PCG with the V-cycle as preconditioner: solves wA x = fb into fx. The convergence test reads one scalar per
// iteration from the device; the true residual is checked at the end like in the Jacobi solver.
int mg_pcg_solve(mof_ctx* ctx, double tol, int maxIters, int* itersOut, double* relresOut) {
    Multigrid& mg = *ctx->mg;
    const int E = ctx->E;
    const double* b = ctx->fb.p;
    double* x = ctx->fx.p;
    double* r = mg.fr.p;
    double* p = mg.fp.p;
    double* q = mg.fq.p;
    const int fineGrid = kSMs * 8;
    MOF_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * E, ctx->stream));
    MOF_CUDA(cudaMemcpyAsync(r, b, sizeof(double) * E, cudaMemcpyDeviceToDevice, ctx->stream));
    MOF_LAUNCH(k_dot_partial, NBLK, B, 0, b, b, E, mg.partial.p);
    MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_BB, mg.scal.p);
    double bb = 0;
    MOF_CUDA(cudaMemcpyAsync(&bb, mg.scal.p + S_BB, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    *itersOut = 0, *relresOut = 0;
    if (!(bb > 0)) return MOF_OK;
    int it = 0;
    double rr = bb;
    for (int attempt = 0; attempt < 4; attempt++) {
        // (re)start: z = M r, p = z, rz = r.z
        MOF_TRY(fine_cycle(ctx, mg, r));
        MOF_LAUNCH(k_dot_partial, NBLK, B, 0, r, mg.fz.p, E, mg.partial.p);
        MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_RZ, mg.scal.p);
        MOF_LAUNCH(k_direction, blocks_for(E, B), B, 0, mg.fz.p, mg.scal.p, E, 1, p);
        // One PCG iteration is ~50-150 small dependent launches (most of them on the tiny coarse levels): capture
        // TWO iterations once as a CUDA graph and replay it (the ping-pong buffers of the cycle are back in place
        // after an even number of cycles). The residual norms of both iterations land in pinned host memory.
        auto iteration = [&](int slot) -> int {
            int np = 0;
            MOF_TRY(spmv_dot_launch(ctx, E, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, p, q, mg.partial.p, &np));
            MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, np, S_PQ, mg.scal.p);
            MOF_LAUNCH(k_update_xr, NBLK, B, 0, p, q, mg.scal.p, E, x, r, mg.partial.p);
            MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_RR, mg.scal.p);
            MOF_CUDA(cudaMemcpyAsync(mg.hostRR + slot, mg.scal.p + S_RR, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MOF_TRY(fine_cycle(ctx, mg, r));
            MOF_LAUNCH(k_dot_partial, NBLK, B, 0, r, mg.fz.p, E, mg.partial.p);
            MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_RZNEW, mg.scal.p);
            MOF_LAUNCH(k_direction, blocks_for(E, B), B, 0, mg.fz.p, mg.scal.p, E, 0, p);
            return MOF_OK;
        };
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const long long launchesBefore = ctx->stats.kernelLaunches;
        MOF_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
        int crc = iteration(0);
        if (crc == MOF_OK) crc = iteration(1);
        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
        const long long launchesPerReplay = ctx->stats.kernelLaunches - launchesBefore;
        ctx->stats.kernelLaunches = launchesBefore;
        if (crc != MOF_OK || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            return crc != MOF_OK ? crc : cuda_fail(ctx, ce, "cudaStreamEndCapture(mg iteration)");
        }
        ce = cudaGraphInstantiate(&exec, graph, 0);
        if (ce != cudaSuccess) {
            cudaGraphDestroy(graph);
            return cuda_fail(ctx, ce, "cudaGraphInstantiate(mg iteration)");
        }
        while (it < maxIters) {
            ce = cudaGraphLaunch(exec, ctx->stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
            if (ce != cudaSuccess) break;
            ctx->stats.kernelLaunches += launchesPerReplay;
            it += 2;
            rr = mg.hostRR[1];
            if (!(mg.hostRR[0] > tol * tol * bb) || !(rr > tol * tol * bb) || !std::isfinite(rr)) break;
        }
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return cuda_fail(ctx, ce, "cudaGraphLaunch(mg iteration)");
        // true residual of x
        MOF_LAUNCH(k_fine_apply, fineGrid, B, 0, E, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, b, ctx->wDinv.p, 0., x, r, 1);
        MOF_LAUNCH(k_dot_partial, NBLK, B, 0, r, r, E, mg.partial.p);
        MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_RR, mg.scal.p);
        MOF_CUDA(cudaMemcpyAsync(&rr, mg.scal.p + S_RR, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MOF_CUDA(cudaStreamSynchronize(ctx->stream));
        if (!(rr > tol * tol * bb * 1.0002) || it >= maxIters || !std::isfinite(rr)) break;
    }
    *itersOut = it;
    *relresOut = std::sqrt(rr / bb);
    if (!std::isfinite(rr) || (!(*relresOut <= tol * 1.0001) && (it >= maxIters || !(*relresOut <= 1e-4)))) {
        char msg[160];
        snprintf(msg, sizeof(msg), "[ERROR] multigrid PCG did not reach %g in %d iterations (relative residual %g)", tol, it, *relresOut);
        return fail(ctx, MOF_E_NOCONVERGE, msg);
    }
    return MOF_OK;
}

}  // namespace mof
