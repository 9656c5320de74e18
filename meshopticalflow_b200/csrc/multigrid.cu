// Multilevel preconditioners for the two families of linear systems of the alignment loop, used inside PCG in
// place of the plain Jacobi scaling:
//   FLOW    A = s*R D P + w*S on the E Whitney (edge) unknowns          (VectorField.h:67; one right-hand side)
//   SCALAR  M + eps*S on the V vertices, six right-hand sides at once   (OpticalFlow.cpp:355, :828)
// The reference factorises both exactly (Eigen SimplicialLDLT / LLT). An iterative solve with Jacobi needs
// ~sqrt(cond) iterations (4 500 for the flow system at 3.1M unknowns, 1 900 for the first smoothing solve at 1M
// vertices); one multigrid cycle per iteration brings that to ~70 and ~15.
//
// Construction (aggregation multigrid on an octree; built on the GPU, deterministic):
//  * Aggregates are the occupied cells of a uniform grid over the unknowns' positions (edge midpoints / vertices);
//    coarser levels are the parent cells of the octree, down to <= 64 cells where a dense inverse is applied.
//    Cells are numbered in Morton order, so the children of a cell are contiguous.
//  * Coarse space. SCALAR: one unknown per cell (piecewise constants), applied to the six channels alike.
//    FLOW: a smooth tangent field looks locally like a constant vector c of the ambient space, whose Whitney
//    coefficient on an edge is c . (x_head - x_tail); so every cell carries three unknowns (c_x, c_y, c_z) and
//    the prolongation of an edge is its edge vector.
//  * Cells are at least as wide as the longest edge, so coupled aggregates are grid neighbours and every coarse
//    operator is a 27-point stencil (of 3x3 blocks / of scalars): no sparse pattern, no SpGEMM. The Galerkin
//    coefficients are re-summed from the current matrix once per system (they follow the data term / eps); the
//    octree, neighbour tables and per-entry stencil slots depend on the mesh only.
//  * (1,1) cycle — W on the large levels, V on the small ones, see Multigrid::gamma — with damped Jacobi smoothing
//    (block 3x3 pseudo-inverse on the coarse FLOW levels), damping 1.4 / rho with rho from a power iteration per
//    level, which keeps the cycle symmetric positive definite so that plain PCG applies. One PCG iteration is 40-70
//    small dependent launches: two iterations are captured as a CUDA graph and replayed.
// If a mesh does not fit the scheme, or a solve stalls, the callers fall back to the Jacobi-PCG kernel of
// pcg_kernels.cu.
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "mof_internal.cuh"

// Every kernel of this file starts with pdl_wait(), so every launch of this file may be a programmatic dependent launch.
#ifndef MOF_HOST_EMULATION
#undef MOF_LAUNCH
#define MOF_LAUNCH MOF_LAUNCH_PDL
#endif

namespace mof {

namespace {

constexpr int B = 256;
constexpr int NBLK = kSMs * 8;   // fixed grid of the reduction-producing kernels (deterministic partials): full occupancy at 256 threads
constexpr int MAXL = 9;          // finest admissible grid level (512^3 cells)
constexpr int SLOT_CENTER = 13;
constexpr int COARSEST_CELLS = 64;

// Coarse operators are stored component-major: component k of the coefficient (I, slot) at ((slot*K + k) * N + I)
// (K = 9 for FLOW, 1 for SCALAR), so that 32 consecutive cells read 32 consecutive words.
template <int K>
__host__ __device__ __forceinline__ size_t blk(int N, int I, int slot, int k) { return ((size_t)slot * K + k) * (size_t)N + I; }

// The cycle is a preconditioner: it only has to be a fixed symmetric positive definite operator close to the inverse,
// so it runs in single precision (matrix copies, level vectors) and moves half the bytes; the PCG recurrence around
// it, its SpMV, dot products and the true-residual check stay in fp64. The Galerkin sums and block inverses are
// computed in fp64 and rounded once.
#ifdef MOF_MG_FP64
using creal = double;
#else
using creal = float;
#endif

struct MgLevel {
    int gridLevel = 0, N = 0;
    DBuf<int> code, nbr, parent, firstChild;   // firstChild: children (ids on the finer level) of each node, N+1 entries
    DBuf<double> blocks;                       // [27][K][N], setup precision
    DBuf<creal> cblocks, binv;                 // cycle copies: [27][K][N], [K][N]
    DBuf<creal> r, z, t;                       // [N][D]
    double omega = 0.6;
};

}  // namespace

enum MgKind { MG_FLOW = 0, MG_SCALAR = 1 };
// One instantiated CUDA graph of a whole PCG solve (first cycle, a WHILE node around two iterations whose condition a kernel
// of the loop sets on the device, true residual, read-back): launched once per solve, no host in the loop. Nothing the
// next system changes is in its kernel parameters (damping factors, tolerances and counters are read from device memory),
// so it is captured once per mesh and pair of buffers.
struct PcgGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const double* b = nullptr;
    double* x = nullptr;
    long long launchesOnce = 0, launchesPerPass = 0;  // kernels of ours outside the loop / per pass of its body (two iterations)
};
enum { OM_MINUS_ONE = 14, OM_ZERO = 15, OM_COUNT = 16 };  // slots of Multigrid::domega holding -1 (power iteration: z + Minv A z) and 0

struct Multigrid {
    MgKind kind = MG_FLOW;
    bool usable = false;
    int nFine = 0;                  // fine unknowns per right-hand side (E or V)
    int nrhs = 1;                   // 1 (FLOW) or 6 (SCALAR, channels interleaved per vertex)
    int K = 0;                      // number of coarse levels
    std::vector<MgLevel> lev;       // lev[0] = level 1 (finest aggregates)
    DBuf<double> evec;              // FLOW: edge vectors [E][3]
    DBuf<creal> cevec;              // ... in cycle precision
    DBuf<creal> cevecAgg;           // ... and in the order of the aggregates' member lists (aggList): the restriction streams it
    DBuf<int> agg, aggPtr, aggList; // aggregate of every fine unknown; members of every aggregate, ascending
    DBuf<signed char> slotOf;       // per matrix entry: stencil slot at level 1 (-1 for padding)
    DBuf<creal> cinv;               // dense inverse on the coarsest level
    DBuf<creal> fval, fdinv;        // fine matrix values (layout of wA / sSys) and inverse diagonal in cycle precision
    DBuf<creal> fvalSell;           // SCALAR: the same values in the sliced layout of ctx->sSysSell (what the solver's kernels read)
    DBuf<creal> fz, fz2, ft;        // fine-level vectors of the cycle, nFine * nrhs each
    DBuf<double> fr, fp, fq;        // ... and of PCG
    DBuf<double> partial, scal;
    DBuf<unsigned> counter;         // arrival counter of the last-CTA folds (self-resetting)
    double omega0 = 0.6;
    // Damping factors as the kernels read them (device): [0] fine level, [1 + l] coarse level l, two constants. By pointer
    // rather than by value so that a captured PCG graph stays valid when the next system changes them.
    DBuf<double> domega;
    DBuf<double> chebD, chebR, chebC;  // SCALAR: vectors of mg_scalar_cheb
    DBuf<creal> eig;                // power-iteration iterates of the last system: fine level, then the coarse levels
    bool eigValid = false;
    const double* om(int slot) const { return domega.p + slot; }
    double hostOmega[16];
    // Cycle shape: `gamma` coarse corrections per visit on the first `gammaLevels` coarse levels, one below. Default
    // (gammaLevels < 0): W (gamma 2) on all but the four (FLOW) / five (SCALAR) coarsest levels — the large levels,
    // where a second visit halves what PCG has left to do (flow solve at 1M vertices: 68 iterations instead of 108,
    // 4M: 78 instead of 164, 16.8M: 88 instead of 250) at the cost of kernels that are still bandwidth-bound — and V
    // on the small latency-bound levels, where a second visit costs as much as on a large one. Measured per level
    // count; the smoothing systems are better conditioned (most of their solves take ~15 iterations) and want one
    // W level less.
    int gamma = 2, gammaLevels = -1;
    // One mesh over several GPUs: the first distLevels coarse levels are dealt to the ranks in cell ranges (cellStart[l], world + 1
    // entries, a rank's range on level l being the children of its range on level l + 1), the rest is replicated. Halo lists
    // (dist.cu partitions): levelPart[l] = cells of level l outside the own range that the own cells' stencils — and, on level 0, the
    // own fine rows' aggregates — read; memberPart = fine rows outside the own row block that belong to the own level-0 cells.
    int distLevels = 0;
    std::vector<std::vector<int>> cellStart;
    std::vector<int> levelPart;
    int memberPart = -1, gatherPart = -1;
    int coarseSweeps[MAXL + 1];     // damped-Jacobi sweeps before / after the correction per coarse level (MOF_MG_COARSE_SWEEPS[_SCALAR]="a,b,...", default 1)
    int fullDepth = 0;              // levels of the hierarchy before any are skipped (MOF_MG_SKIP_CELLS): the W levels are counted on it
    int fineSweeps = 1;             // damped-Jacobi sweeps on the fine level before and after the coarse correction (MOF_MG_FINE_SWEEPS[_SCALAR])
    double* hostRR = nullptr;       // pinned (a slice of ctx->pinned): what a solve reports back (PCG scalars)
    // The small levels as one kernel on one cluster (k_coarse_tail): first level handled there (-1: none), CTAs of the cluster.
    int tailStart = -1, tailGrid = 0;
    bool noWhile = false;           // the driver refused the conditional graph node once: host-driven replays from then on
    DBuf<unsigned long long> tailTrace;
    unsigned char traceOps[256];    // the program of the last launch (what the trace's intervals are)
    int traceN = 0;
    // PCG loops captured once per (right-hand side, solution) pair of buffers and kept: see PcgGraph.
    std::vector<struct PcgGraph> graphs;
    std::vector<double> hostBlocks, hostDense;
    std::vector<int> hostNbr;
    int comps() const { return kind == MG_FLOW ? 9 : 1; }     // K
    int dofs() const { return kind == MG_FLOW ? 3 : 6; }      // D: values per coarse cell
    size_t fineLen() const { return (size_t)nFine * nrhs; }
};

namespace {

// ---------------------------------------------------------------------------------------- utilities

int env_int(const char* name, int fallback) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : fallback;
}

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

// Edge vectors, midpoints and squared lengths of the Whitney unknowns.
__global__ void k_edge_geometry(const double* __restrict__ pos, const int* __restrict__ tri, const int* __restrict__ expanded, int E, double* __restrict__ evec,
                                double* __restrict__ emid, double* __restrict__ len2) { pdl_wait();
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
__global__ void k_minmax_partial(const double* __restrict__ in, long long n, int stride, int offset, int mode, double* __restrict__ partial) { pdl_wait();
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
__global__ void k_minmax_final(const double* __restrict__ partial, int np, int mode, double* __restrict__ out) { pdl_wait();
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
        int q = (int)((p[k] - gm.lo[k]) * gm.inv * s);
        c[k] = (unsigned)min(max(q, 0), (1 << L) - 1);
    }
    return morton3(c[0], c[1], c[2]);
}
__global__ void k_mark_cells(GridMap gm, const double* __restrict__ pts, int n, int L, int* __restrict__ occ) { pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) occ[cell_code(gm, pts + 3 * (size_t)i, L)] = 1;
}
__global__ void k_mark_parents(const int* __restrict__ occ, long long cells, int* __restrict__ occParent) { pdl_wait();
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cells && occ[c]) occParent[c >> 3] = 1;
}
__global__ void k_node_codes(const int* __restrict__ occ, const int* __restrict__ rank, long long cells, int* __restrict__ code) { pdl_wait();
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cells && occ[c]) code[rank[c]] = (int)c;
}
__global__ void k_neighbours(const int* __restrict__ code, const int* __restrict__ occ, const int* __restrict__ rank, int N, int L, int* __restrict__ nbr) { pdl_wait();
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
// `shift` = 3 x (grid levels between a level and the next coarser one of the hierarchy: 1, or 2 where a level is skipped)
__global__ void k_parents(const int* __restrict__ code, const int* __restrict__ rankParent, int N, int Nparent, int shift, int* __restrict__ parent,
                          int* __restrict__ firstChild) { pdl_wait();
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I > N) return;
    if (I == N) { firstChild[Nparent] = N; return; }
    int p = rankParent[(unsigned)code[I] >> shift];
    parent[I] = p;
    if (I == 0 || rankParent[(unsigned)code[I - 1] >> shift] != p) firstChild[p] = I;
}
__global__ void k_point_aggregate(GridMap gm, const double* __restrict__ pts, const int* __restrict__ rank, int n, int L, int* __restrict__ agg, int* __restrict__ count) { pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = rank[cell_code(gm, pts + 3 * (size_t)i, L)];
    agg[i] = a;
    atomicAdd(&count[a], 1);
}
__global__ void k_aggregate_fill(const int* __restrict__ agg, int n, int* cursor, int* __restrict__ list) { pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[atomicAdd(&cursor[agg[i]], 1)] = i;
}
// Ascending member ids within every aggregate (fixed summation order). One warp per aggregate, rank sort.
__global__ void k_aggregate_sort(const int* __restrict__ ptr, int N, const int* __restrict__ in, int* __restrict__ out) { pdl_wait();
    int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (I >= N) return;
    int b = ptr[I], n = ptr[I + 1] - b;
    for (int i = lane; i < n; i += 32) {
        int v = in[b + i], r = 0;
        for (int k = 0; k < n; k++) r += in[b + k] < v;
        out[b + r] = v;
    }
}

__device__ __forceinline__ signed char slot_between(unsigned ci, unsigned cf, int* flags) {
    int dx = (int)compact3(cf) - (int)compact3(ci), dy = (int)compact3(cf >> 1) - (int)compact3(ci >> 1), dz = (int)compact3(cf >> 2) - (int)compact3(ci >> 2);
    if (dx < -1 || dx > 1 || dy < -1 || dy > 1 || dz < -1 || dz > 1) {
        flags[0] = 1;  // coupled unknowns whose cells are not neighbours: the stencil scheme does not apply
        return -1;
    }
    return (signed char)((dx + 1) * 9 + (dy + 1) * 3 + (dz + 1));
}
// Stencil slot of every entry (e,f) of the sliced FLOW matrix.
__global__ void k_entry_slots_flow(const int* __restrict__ wRowptr, const int* __restrict__ sliceBase, const int* __restrict__ wCol, const int* __restrict__ agg,
                                   const int* __restrict__ code, int E, int slices, signed char* __restrict__ slotOf, int* __restrict__ flags) { pdl_wait();
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 32 * slices) return;
    const int longest = (sliceBase[(e >> 5) + 1] - sliceBase[e >> 5]) >> 5;
    int len = e < E ? wRowptr[e + 1] - wRowptr[e] : 0;
    unsigned ci = e < E ? (unsigned)code[agg[e]] : 0;
    for (int j = 0; j < longest; j++) {
        size_t k = sell_pos(sliceBase, e, j);
        slotOf[k] = j < len ? slot_between(ci, (unsigned)code[agg[wCol[k]]], flags) : (signed char)-1;
    }
}
// ... and of every entry (v,u) of the SCALAR CSR pattern.
__global__ void k_entry_slots_scalar(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ agg, const int* __restrict__ code, int V,
                                     signed char* __restrict__ slotOf, int* __restrict__ flags) { pdl_wait();
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    unsigned ci = (unsigned)code[agg[v]];
    for (int k = rowptr[v]; k < rowptr[v + 1]; k++) slotOf[k] = slot_between(ci, (unsigned)code[agg[col[k]]], flags);
}

// ------------------------------------------------------------------------------------ Galerkin values

// FLOW level-1 block (I, slot) = sum over edges e of I, entries f of row e that fall in that neighbour cell, of
// A_ef * v_e v_f^T. One thread per (I, slot); the 27 threads of a cell walk the same entries (broadcast).
__global__ void k_level1_flow(const int* __restrict__ aggPtr, const int* __restrict__ aggList, const int* __restrict__ wRowptr, const int* __restrict__ sliceBase,
                              const int* __restrict__ wCol, const double* __restrict__ wA, const signed char* __restrict__ slotOf, const double* __restrict__ evec,
                              const int* __restrict__ nbr, int N, double scale, double* __restrict__ blocks) { pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 27) return;
    int I = i / 27, s = i - 27 * I;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (nbr[i] >= 0)
        for (int q = aggPtr[I]; q < aggPtr[I + 1]; q++) {
            int e = aggList[q], len = wRowptr[e + 1] - wRowptr[e];
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
    for (int k = 0; k < 9; k++) blocks[blk<9>(N, I, s, k)] = scale * acc[k];
}
// SCALAR level-1 weight (I, slot) = sum over vertices v of I of the entries of row v that fall in that cell.
__global__ void k_level1_scalar(const int* __restrict__ aggPtr, const int* __restrict__ aggList, const int* __restrict__ rowptr, const double* __restrict__ val,
                                const signed char* __restrict__ slotOf, const int* __restrict__ nbr, int N, double scale, double* __restrict__ blocks) { pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * 27) return;
    int I = i / 27, s = i - 27 * I;
    double acc = 0;
    if (nbr[i] >= 0)
        for (int q = aggPtr[I]; q < aggPtr[I + 1]; q++) {
            int v = aggList[q];
            for (int k = rowptr[v]; k < rowptr[v + 1]; k++)
                if (slotOf[k] == s) acc += val[k];
        }
    blocks[blk<1>(N, I, s, 0)] = scale * acc;
}

// Coarser coefficient (I', slot') = sum of the finer ones (I, s) with parent(I) = I' and parent(nbr(I, s)) = nbr'(I', slot').
template <int K>
__global__ void k_coarsen_blocks(const int* __restrict__ firstChild, const int* __restrict__ nbrFine, const int* __restrict__ parentFine, const double* __restrict__ blocksFine,
                                 int Nfine, const int* __restrict__ nbrCoarse, int Ncoarse, double scale, double* __restrict__ blocksCoarse) { pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ncoarse * 27) return;
    int Ip = i / 27, sp = i - 27 * Ip;
    int Jp = nbrCoarse[i];
    double acc[K];
    for (int k = 0; k < K; k++) acc[k] = 0;
    if (Jp >= 0)
        for (int I = firstChild[Ip]; I < firstChild[Ip + 1]; I++)
            for (int s = 0; s < 27; s++) {
                int J = nbrFine[I * 27 + s];
                if (J < 0 || parentFine[J] != Jp) continue;
                for (int k = 0; k < K; k++) acc[k] += blocksFine[blk<K>(Nfine, I, s, k)];
            }
    for (int k = 0; k < K; k++) blocksCoarse[blk<K>(Ncoarse, Ip, sp, k)] = scale * acc[k];
}

// FLOW: pseudo-inverse of the diagonal 3x3 block through its eigen-decomposition (cyclic Jacobi rotations, accurate
// for each eigenvalue separately). Directions with eigenvalue < relTol * largest are not smoothed on that level:
//  * cells on the fringe of the surface hold one or two edges, or coplanar ones, and their block is rank deficient;
//  * inside a cell the surface is nearly flat, so the direction normal to it carries an eigenvalue ~(cell size x
//    curvature)^2 of the tangential ones (2e-4 at 1M vertices, 1.5e-5 at 16M). Keeping it (relTol 1e-8) makes block
//    Jacobi badly scaled on large meshes: at 16.8M vertices rho(Binv A) = 4.0 on level 1 and the solve stalls;
//    with relTol = 1e-3 rho stays at 1.8 on every level and size (and 4M vertices need 164 iterations, not 194).
// An explicit cofactor inverse of such a block is NOT good enough: its error in the well-conditioned directions scales
// with the condition number and the smoother stops being positive definite.
__global__ void k_block_pinv(const double* __restrict__ blocks, int N, double relTol, creal* __restrict__ binv) { pdl_wait();
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= N) return;
    double m[9];
    for (int k = 0; k < 9; k++) m[k] = blocks[blk<9>(N, I, SLOT_CENTER, k)];
    double a00 = m[0], a11 = m[4], a22 = m[8], a01 = 0.5 * (m[1] + m[3]), a02 = 0.5 * (m[2] + m[6]), a12 = 0.5 * (m[5] + m[7]);
    double Vm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // columns = eigenvectors
    for (int sweep = 0; sweep < 8; sweep++) {
#define MOF_ROT(app, aqq, apq, arp, arq, p, q)                                           \
    if (apq != 0.) {                                                                     \
        double theta = (aqq - app) / (2. * apq);                                         \
        double t = (theta >= 0 ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1.));   \
        double c = 1. / sqrt(t * t + 1.), s = t * c;                                     \
        double npp = app - t * apq, nqq = aqq + t * apq;                                 \
        double nrp = c * arp - s * arq, nrq = s * arp + c * arq;                         \
        app = npp, aqq = nqq, apq = 0., arp = nrp, arq = nrq;                            \
        for (int r = 0; r < 3; r++) {                                                    \
            double vp = Vm[3 * r + p], vq = Vm[3 * r + q];                               \
            Vm[3 * r + p] = c * vp - s * vq, Vm[3 * r + q] = s * vp + c * vq;            \
        }                                                                                \
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
        if (!(lam[e] > relTol * top) || !(top > 0)) continue;
        double il = 1. / lam[e];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) inv[3 * r + c] += il * Vm[3 * r + e] * Vm[3 * c + e];
    }
    for (int k = 0; k < 9; k++) binv[(size_t)k * N + I] = (creal)inv[k];
}
__global__ void k_scalar_inv(const double* __restrict__ blocks, int N, creal* __restrict__ binv) { pdl_wait();
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= N) return;
    double w = blocks[blk<1>(N, I, SLOT_CENTER, 0)];
    binv[I] = (creal)(w > 0 ? 1. / w : 0.);
}
// fp64 -> cycle precision (matrix values, inverse diagonals, coarse coefficients, edge vectors)
__global__ void k_to_creal(const double* __restrict__ src, long long n, creal* __restrict__ dst) { pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = (creal)src[i];
}

// ------------------------------------------------------------------------------------- cycle kernels

// z = omega * dinv[row] * r, flat over nrhs interleaved right-hand sides
template <class TZ>
__global__ void k_fine_presmooth(const double* __restrict__ r, const creal* __restrict__ dinv, const double* __restrict__ omegaP, long long len, int nrhs,
                                 TZ* __restrict__ z) { pdl_wait();
    const double omega = *omegaP;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) z[i] = (TZ)(omega * (double)dinv[i / nrhs] * r[i]);
}

// Slots of the PCG scalars (mg.scal).
enum { S_RZ = 0, S_PQ = 1, S_ALPHA = 2, S_BETA = 3, S_RR = 4, S_BB = 5, S_RZNEW = 6,
       // the device-side loop control of a captured solve (slots 16..39 belong to the power iterations of the set-up)
       S_THR = 40,    // tol^2 * b.b: the recurrence residual r.r stops the loop at or below it
       S_IT = 41,     // iterations done
       S_MAXIT = 42,  // ... and their limit
       S_RR0 = 43,    // r.r after the first of the two iterations of a pass
       S_REPORT = 48 };  // scal[0 .. S_REPORT) is what a solve copies back to the host

// One CTA (all B threads): adds `np` per-CTA partials in index order into scal[slot] and derives the PCG scalar that
// depends on it.
__device__ __forceinline__ void fold_partials(const double* partial, int np, int slot, double* scal) {
    __shared__ double shf[B];
    double s = 0;
    for (int i = threadIdx.x; i < np; i += B) s += __ldcg(partial + i);
    shf[threadIdx.x] = s;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) shf[threadIdx.x] += shf[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double v = shf[0];
        scal[slot] = v;
        if (slot == S_PQ) scal[S_ALPHA] = v != 0 ? scal[S_RZ] / v : 0.;
        if (slot == S_RZNEW) {
            scal[S_BETA] = scal[S_RZ] != 0 ? v / scal[S_RZ] : 0.;
            scal[S_RZ] = v;
        }
    }
}
// Sum of one value per thread over the CTA, in a fixed order; thread 0 stores it in partial[blockIdx.x]. With slot >= 0
// the last CTA of the grid to arrive (a self-resetting counter) also folds all the partials — in index order, so the
// result does not depend on which CTA that is — which saves a one-CTA launch on the critical path.
__device__ __forceinline__ void cta_partial(double v, double* partial, int slot = -1, double* scal = nullptr, unsigned* counter = nullptr) {
    __shared__ double sh[B];
    __shared__ bool last;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = sh[0];
        if (slot >= 0) {
            __threadfence();
            last = atomicAdd(counter, 1u) == gridDim.x - 1;
        }
    }
    if (slot < 0) return;
    __syncthreads();
    if (!last) return;
    __threadfence();
    fold_partials(partial, (int)gridDim.x, slot, scal);
    if (threadIdx.x == 0) *counter = 0;
}

// Where a kernel's reduction goes: per-CTA partials (nullptr: none) and, with slot >= 0, the fold by the last CTA.
struct Fold {
    double* partial;
    int slot;
    double* scal;
    unsigned* counter;
};

// FLOW: sliced SpMV (warp = slice, lane = row), as in pcg_kernels.cu. TV = precision of the matrix copy, TX = of the
// vectors; b is always the fp64 right-hand side (the PCG residual inside a cycle).
//   mode 1: out = b - A in
//   mode 2: out = in + omega * dinv * (b - A in)  (one damped Jacobi sweep); with `partial`, also the CTA's part of b.out
// PART (one mesh over several GPUs, dist.cu): only the slices [s0, s1) of this rank, and
//   mode 0: out = A in with the CTA's part of in.out
constexpr int BATCH = 6;
template <class TV, class TX, bool PART = false>
__global__ void __launch_bounds__(B, 8) k_fine_apply_flow(int n, const int* __restrict__ sliceBase, const int* __restrict__ col, const TV* __restrict__ val,
                                                      const double* __restrict__ b, const creal* __restrict__ dinv, const double* __restrict__ omegaP,
                                                      const TX* __restrict__ in, TX* __restrict__ out, int mode, Fold f, int s0 = 0, int s1 = 0) { pdl_wait();
    const double omega = *omegaP;
    const int lane = threadIdx.x & 31;
    const int slices = PART ? s1 : (n + 31) >> 5;
    const int warps = gridDim.x * (B / 32);
    double dot = 0;
    for (int s = (PART ? s0 : 0) + blockIdx.x * (B / 32) + (threadIdx.x >> 5); s < slices; s += warps) {
        const int base = sliceBase[s];
        const int len = (sliceBase[s + 1] - base) >> 5;
        const TV* v0 = val + (size_t)base + lane;
        const int* c0 = col + (size_t)base + lane;
        const int row = 32 * s + lane;
        TX acc = 0;
        for (int j0 = 0; j0 < len; j0 += BATCH) {
            TV v[BATCH];
            TX x[BATCH];
            int c[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; u++) {
                bool ok = j0 + u < len;
                v[u] = ok ? __ldcs(v0 + 32 * (size_t)(j0 + u)) : (TV)0;
                c[u] = ok ? __ldcs(c0 + 32 * (size_t)(j0 + u)) : 0;
            }
#pragma unroll
            for (int u = 0; u < BATCH; u++) x[u] = in[c[u]];
#pragma unroll
            for (int u = 0; u < BATCH; u++)
                if (j0 + u < len) acc += (TX)v[u] * x[u];
        }
        if (row < n) {
            if (PART && mode == 0) {
                out[row] = acc;
                dot += (double)in[row] * (double)acc;
                continue;
            }
            const double bv = b[row];
            if (mode == 1) out[row] = (TX)(bv - (double)acc);
            else {
                TX o = (TX)((double)in[row] + omega * (double)dinv[row] * (bv - (double)acc));
                out[row] = o;
                dot += bv * (double)o;
            }
        }
    }
    if ((mode == 2 || (PART && mode == 0)) && f.partial) cta_partial(dot, f.partial, f.slot, f.scal, f.counter);
}
// SCALAR: CSR with six interleaved right-hand sides, one thread per ROW, the six channels of a vertex in registers,
// gathered and stored as three 8- or 16-byte pieces (six times fewer val/col/rowptr loads than one thread per
// (row, channel); 6 % faster smoothing solves at 1M vertices). Same modes as k_fine_apply_scalar below, which remains
// for the row-partitioned path and under MOF_SCALAR_ROWKERNEL=0.
template <class T> struct Pair;
template <> struct Pair<float> { using type = float2; };
template <> struct Pair<double> { using type = double2; };
template <class TV, class TX>
__global__ void __launch_bounds__(B) k_fine_apply_scalar_row(int n, const int* __restrict__ rowptr, const int* __restrict__ col, const TV* __restrict__ val,
                                                            const double* __restrict__ b, const creal* __restrict__ dinv, const double* __restrict__ omegaP,
                                                            const TX* __restrict__ in, TX* __restrict__ out, int mode, Fold f) { pdl_wait();
    const double omega = *omegaP;
    using P2 = typename Pair<TX>::type;
    double dot = 0;
    for (int row = blockIdx.x * B + threadIdx.x; row < n; row += gridDim.x * B) {
        const int k0 = rowptr[row], k1 = rowptr[row + 1];
        TX acc[6] = {0, 0, 0, 0, 0, 0};
        for (int k = k0; k < k1; k++) {
            const TX v = (TX)val[k];
            const P2* src = reinterpret_cast<const P2*>(in + 6 * (size_t)col[k]);
            const P2 x0 = src[0], x1 = src[1], x2 = src[2];
            acc[0] += v * x0.x, acc[1] += v * x0.y, acc[2] += v * x1.x, acc[3] += v * x1.y, acc[4] += v * x2.x, acc[5] += v * x2.y;
        }
        const P2* self = reinterpret_cast<const P2*>(in + 6 * (size_t)row);
        TX o[6];
        if (mode == 0) {
            const P2 s0 = self[0], s1 = self[1], s2 = self[2];
            const TX sv[6] = {s0.x, s0.y, s1.x, s1.y, s2.x, s2.y};
#pragma unroll
            for (int c = 0; c < 6; c++) o[c] = acc[c], dot += (double)sv[c] * (double)acc[c];
        } else {
            const double2* bp = reinterpret_cast<const double2*>(b + 6 * (size_t)row);
            const double2 b0 = bp[0], b1 = bp[1], b2 = bp[2];
            const double bv[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
            if (mode == 1) {
#pragma unroll
                for (int c = 0; c < 6; c++) o[c] = (TX)(bv[c] - (double)acc[c]);
            } else {
                const P2 s0 = self[0], s1 = self[1], s2 = self[2];
                const TX sv[6] = {s0.x, s0.y, s1.x, s1.y, s2.x, s2.y};
                const double wd = omega * (double)dinv[row];
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    o[c] = (TX)((double)sv[c] + wd * (bv[c] - (double)acc[c]));
                    dot += bv[c] * (double)o[c];
                }
            }
        }
        P2* dst = reinterpret_cast<P2*>(out + 6 * (size_t)row);
        P2 w0, w1, w2;
        w0.x = o[0], w0.y = o[1], w1.x = o[2], w1.y = o[3], w2.x = o[4], w2.y = o[5];
        dst[0] = w0, dst[1] = w1, dst[2] = w2;
    }
    if (mode == 0 || (mode == 2 && f.partial)) cta_partial(dot, f.partial, f.slot, f.scal, f.counter);
}
// SCALAR, sliced layout (SELL-32, like the E x E operators): warp = slice of 32 rows, lane = row, the six channels of the row in
// registers. Entry j of the 32 rows is 32 consecutive words, so every val / col load is one coalesced request, and the slice's
// common length lets the loads of a row be issued in batches ahead of the gathers they feed (the CSR row kernel above walks
// its row entry by entry: rowptr -> col -> gather, one dependent chain per entry; 3.0-3.7 TB/s against 5-6 here). The entries of
// a row keep their CSR order, so the sums are the same bits. Modes as k_fine_apply_scalar_row.
constexpr int SBATCH = 4;
template <class TV, class TX>
__global__ void __launch_bounds__(B, 4) k_fine_apply_scalar_sell(int n, const int* __restrict__ sliceBase, const int* __restrict__ col, const TV* __restrict__ val,
                                                                const double* __restrict__ b, const creal* __restrict__ dinv, const double* __restrict__ omegaP,
                                                                const TX* __restrict__ in, TX* __restrict__ out, int mode, Fold f) { pdl_wait();
    using P2 = typename Pair<TX>::type;
    const double omega = *omegaP;
    const int lane = threadIdx.x & 31;
    const int slices = (n + 31) >> 5;
    const int warps = gridDim.x * (B / 32);
    double dot = 0;
    for (int s = blockIdx.x * (B / 32) + (threadIdx.x >> 5); s < slices; s += warps) {
        const int base = sliceBase[s];
        const int len = (sliceBase[s + 1] - base) >> 5;
        const TV* v0 = val + (size_t)base + lane;
        const int* c0 = col + (size_t)base + lane;
        const int row = 32 * s + lane;
        TX acc[6] = {0, 0, 0, 0, 0, 0};
        for (int j0 = 0; j0 < len; j0 += SBATCH) {
            TV v[SBATCH];
            int c[SBATCH];
            P2 x[SBATCH][3];
#pragma unroll
            for (int u = 0; u < SBATCH; u++) {
                const bool ok = j0 + u < len;
                v[u] = ok ? __ldcs(v0 + 32 * (size_t)(j0 + u)) : (TV)0;
                c[u] = ok ? __ldcs(c0 + 32 * (size_t)(j0 + u)) : 0;
            }
#pragma unroll
            for (int u = 0; u < SBATCH; u++) {
                const P2* src = reinterpret_cast<const P2*>(in + 6 * (size_t)c[u]);
                x[u][0] = src[0], x[u][1] = src[1], x[u][2] = src[2];
            }
#pragma unroll
            for (int u = 0; u < SBATCH; u++)
                if (j0 + u < len) {
                    const TX w = (TX)v[u];
                    acc[0] += w * x[u][0].x, acc[1] += w * x[u][0].y, acc[2] += w * x[u][1].x, acc[3] += w * x[u][1].y, acc[4] += w * x[u][2].x, acc[5] += w * x[u][2].y;
                }
        }
        if (row >= n) continue;
        const P2* self = reinterpret_cast<const P2*>(in + 6 * (size_t)row);
        TX o[6];
        if (mode == 0) {
            const P2 s0 = self[0], s1 = self[1], s2 = self[2];
            const TX sv[6] = {s0.x, s0.y, s1.x, s1.y, s2.x, s2.y};
#pragma unroll
            for (int c2 = 0; c2 < 6; c2++) o[c2] = acc[c2], dot += (double)sv[c2] * (double)acc[c2];
        } else {
            const double2* bp = reinterpret_cast<const double2*>(b + 6 * (size_t)row);
            const double2 b0 = bp[0], b1 = bp[1], b2 = bp[2];
            const double bv[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
            if (mode == 1) {
#pragma unroll
                for (int c2 = 0; c2 < 6; c2++) o[c2] = (TX)(bv[c2] - (double)acc[c2]);
            } else {
                const P2 s0 = self[0], s1 = self[1], s2 = self[2];
                const TX sv[6] = {s0.x, s0.y, s1.x, s1.y, s2.x, s2.y};
                const double wd = omega * (double)dinv[row];
#pragma unroll
                for (int c2 = 0; c2 < 6; c2++) {
                    o[c2] = (TX)((double)sv[c2] + wd * (bv[c2] - (double)acc[c2]));
                    dot += bv[c2] * (double)o[c2];
                }
            }
        }
        P2* dst = reinterpret_cast<P2*>(out + 6 * (size_t)row);
        P2 w0, w1, w2;
        w0.x = o[0], w0.y = o[1], w1.x = o[2], w1.y = o[3], w2.x = o[4], w2.y = o[5];
        dst[0] = w0, dst[1] = w1, dst[2] = w2;
    }
    if (mode == 0 || (mode == 2 && f.partial)) cta_partial(dot, f.partial, f.slot, f.scal, f.counter);
}
// CSR values -> the sliced layout, in both precisions (once per scalar system; the pattern's slice offsets are per mesh).
__global__ void k_scalar_vals_to_sell(const int* __restrict__ rowptr, const double* __restrict__ val, const int* __restrict__ sliceBase, int n, int slices,
                                      double* __restrict__ sVal, creal* __restrict__ cVal) { pdl_wait();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= 32 * slices) return;
    const int longest = (sliceBase[(r >> 5) + 1] - sliceBase[r >> 5]) >> 5;
    const int k0 = r < n ? rowptr[r] : 0, len = r < n ? rowptr[r + 1] - k0 : 0;
    for (int j = 0; j < longest; j++) {
        const double v = j < len ? val[k0 + j] : 0.;
        const size_t p = sell_pos(sliceBase, r, j);
        sVal[p] = v, cVal[p] = (creal)v;
    }
}

// SCALAR: CSR with six interleaved right-hand sides, one thread per (row, channel). mode 0: out = A in with the CTA's
// partial of in.out ; modes 1, 2 as above.
// PART (one mesh over several GPUs): only the rows [r0, r1) of this rank.
template <class TV, class TX, bool PART = false>
__global__ void __launch_bounds__(B, 8) k_fine_apply_scalar(int n, const int* __restrict__ rowptr, const int* __restrict__ col, const TV* __restrict__ val,
                                                        const double* __restrict__ b, const creal* __restrict__ dinv, const double* __restrict__ omegaP,
                                                        const TX* __restrict__ in, TX* __restrict__ out, int mode, Fold f, int r0 = 0, int r1 = 0) { pdl_wait();
    const double omega = *omegaP;
    const long long len = PART ? 6ll * r1 : 6ll * n;
    double dot = 0;
    for (long long i = (PART ? 6ll * r0 : 0ll) + (long long)blockIdx.x * B + threadIdx.x; i < len; i += (long long)gridDim.x * B) {
        const int row = (int)(i / 6), c = (int)(i - 6ll * row);
        TX acc = 0;
        // (an explicit batch of 8 predicated loads per row was measured: 86 us against 57 us for this plain loop)
        for (int k = rowptr[row]; k < rowptr[row + 1]; k++) acc += (TX)val[k] * in[6 * (size_t)col[k] + c];
        if (mode == 0) out[i] = acc, dot += (double)in[i] * (double)acc;
        else if (mode == 1) out[i] = (TX)(b[i] - (double)acc);
        else {
            const double bv = b[i];
            TX o = (TX)((double)in[i] + omega * (double)dinv[row] * (bv - (double)acc));
            out[i] = o;
            dot += bv * (double)o;
        }
    }
    if (mode == 0 || (mode == 2 && f.partial)) cta_partial(dot, f.partial, f.slot, f.scal, f.counter);
}

__device__ __forceinline__ void mat3_vec(const creal* m, const creal* v, creal* out) {
    out[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    out[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    out[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
// Applies the diagonal (pseudo-)inverse of cell I to the D values in `v`.
template <int K, int D>
__device__ __forceinline__ void apply_binv(const creal* __restrict__ binv, int N, int I, const creal* v, creal* out) {
    if (K == 9) {
        creal m[9];
#pragma unroll
        for (int k = 0; k < 9; k++) m[k] = binv[(size_t)k * N + I];
        mat3_vec(m, v, out);
    } else {
        creal w = binv[I];
#pragma unroll
        for (int c = 0; c < D; c++) out[c] = w * v[c];
    }
}

// The edge vectors in member-list order: evecAgg[q] = evec[aggList[q]] (once per mesh).
__global__ void k_gather_evec(const int* __restrict__ aggList, const creal* __restrict__ evec, int n, creal* __restrict__ out) { pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3ll * n) return;
    const long long q = i / 3;
    out[i] = evec[3 * (size_t)aggList[q] + (i - 3 * q)];
}
// Every restriction also performs the first smoothing sweep of the level it lands on (from a zero guess):
// zc = omega * Binv * rc.
// FLOW restriction: rc[I] = sum over the edges of aggregate I of v_e * r_e (P1^T r). One warp per aggregate; `evec` = the edge vectors in
// the order of the member lists (Multigrid::cevecAgg), so that only r is gathered.
__global__ void k_restrict_flow(const int* __restrict__ aggPtr, const int* __restrict__ aggList, const creal* __restrict__ evec, const creal* __restrict__ r, int N,
                                const creal* __restrict__ binv, const double* __restrict__ omegaP, creal* __restrict__ rc, creal* __restrict__ zc, int r0, int r1,
                                int c0 = 0, int c1 = -1) { pdl_wait();  // [c0, c1): the cells of this launch (a rank's own cells when level 1 is dealt to the ranks)
    int I = c0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (I >= (c1 < 0 ? N : c1)) return;
    const creal omega = (creal)*omegaP;
    const creal bk = lane < 9 ? binv[(size_t)lane * N + I] : (creal)0;  // the cell's block inverse, fetched alongside the sums
    creal a0 = 0, a1 = 0, a2 = 0;
    for (int q = aggPtr[I] + lane; q < aggPtr[I + 1]; q += 32) {
        int e = aggList[q];
        if (e < r0 || e >= r1) continue;  // partitioned mesh: this rank's rows only (the partial sums are all-reduced)
        creal re = r[e];
        a0 += evec[3 * (size_t)q] * re, a1 += evec[3 * (size_t)q + 1] * re, a2 += evec[3 * (size_t)q + 2] * re;  // (evec in member-list order: coalesced)
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o), a1 += __shfl_xor_sync(0xffffffffu, a1, o), a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    creal m[9];
#pragma unroll
    for (int k = 0; k < 9; k++) m[k] = __shfl_sync(0xffffffffu, bk, k);
    if (lane == 0) {
        creal v[3] = {a0, a1, a2}, o[3];
        mat3_vec(m, v, o);
        for (int c = 0; c < 3; c++) rc[3 * (size_t)I + c] = v[c], zc[3 * (size_t)I + c] = omega * o[c];
    }
}
// FLOW prolongation: z_e += v_e . zc[agg(e)]
__global__ void k_prolong_flow(const int* __restrict__ agg, const creal* __restrict__ evec, const creal* __restrict__ zc, int E, creal* __restrict__ z) { pdl_wait();
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const creal* c = zc + 3 * (size_t)agg[e];
    z[e] += evec[3 * (size_t)e] * c[0] + evec[3 * (size_t)e + 1] * c[1] + evec[3 * (size_t)e + 2] * c[2];
}
// SCALAR restriction / prolongation: sums and copies per channel. One thread per (cell, channel) / (vertex, channel).
__global__ void k_restrict_scalar(const int* __restrict__ aggPtr, const int* __restrict__ aggList, const creal* __restrict__ r, int N, const creal* __restrict__ binv,
                                  const double* __restrict__ omegaP, creal* __restrict__ rc, creal* __restrict__ zc, int r0, int r1, int c0 = 0, int c1 = -1) { pdl_wait();
    int i = 6 * c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * (c1 < 0 ? N : c1)) return;
    const creal omega = (creal)*omegaP;
    int I = i / 6, c = i - 6 * I;
    creal a = 0;
    for (int q = aggPtr[I]; q < aggPtr[I + 1]; q++) {
        const int v = aggList[q];
        if (v >= r0 && v < r1) a += r[6 * (size_t)v + c];  // (partitioned mesh: this rank's rows only)
    }
    rc[i] = a;
    zc[i] = omega * binv[I] * a;
}
__global__ void k_prolong_scalar(const int* __restrict__ agg, const creal* __restrict__ zc, int V, creal* __restrict__ z) { pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6ll * V) return;
    long long v = i / 6;
    z[i] += zc[6 * (size_t)agg[v] + (i - 6 * v)];
}

// Coarse operator of a level. mode 1: out = r - A z ; mode 2: out = z + omega * Binv (r - A z). With `zc`, z stands for
// z + (prolongation of zc), i.e. the coarse correction is added on the fly while gathering (z itself is not written).
// CTA = 32 consecutive cells x (27 / SLOTS) warps, a warp handling SLOTS stencil slots of those cells (lane = cell): every
// coefficient component is read as 32 consecutive words and all of a thread's loads are issued before the first use
// (a missing neighbour reads the cell itself against an all-zero coefficient, so there is no branch). SLOTS = 3 on the
// large levels: 288-thread CTAs, seven of which overlap on an SM; SLOTS = 1 on the small, latency-bound ones. The
// partial sums per (cell, component) are then added in warp order by 32*D threads, which also apply the block inverse.
template <int K, int D, int SLOTS>
__global__ void __launch_bounds__(27 / SLOTS * 32) k_coarse_apply(const creal* __restrict__ blocks, const int* __restrict__ nbr, const creal* __restrict__ binv,
                                                                 const creal* __restrict__ r, const creal* __restrict__ z, const double* __restrict__ omegaP, int N,
                                                                 int mode, creal* __restrict__ out, const int* __restrict__ parent, const creal* __restrict__ zc,
                                                                 int c0 = 0, int c1 = -1) { pdl_wait();  // [c0, c1): the cells of this launch (all, or a rank's own)
    constexpr int WARPS = 27 / SLOTS;
    const creal omega = (creal)*omegaP;
    __shared__ creal part[WARPS][32 * D];
    __shared__ creal resS[32 * D];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int I = c0 + blockIdx.x * 32 + lane;
    const int cellEnd = c1 < 0 ? N : c1;
    // operands of the finishing step (thread t finishes component c of cell Ic), fetched ahead of the barrier
    const int t = threadIdx.x;
    const int ln = t / D, c = t - D * ln;
    const int Ic = c0 + blockIdx.x * 32 + ln;
    const bool live = t < 32 * D && Ic < cellEnd;
    creal rv = 0, zi = 0, bi[3] = {0, 0, 0};
    if (live) {
        rv = r[(size_t)D * Ic + c];
        if (mode == 2) {
            zi = z[(size_t)D * Ic + c];
            if (zc) zi += zc[(size_t)D * parent[Ic] + c];
            if (K == 9) {
#pragma unroll
                for (int k = 0; k < 3; k++) bi[k] = binv[(size_t)(3 * c + k) * N + Ic];
            } else
                bi[0] = binv[Ic];
        }
    }
    creal o[D];
#pragma unroll
    for (int c2 = 0; c2 < D; c2++) o[c2] = 0;
    if (I < cellEnd) {
        int J[SLOTS], P[SLOTS];
        creal m[SLOTS][K], zj[SLOTS][D];
#pragma unroll
        for (int u = 0; u < SLOTS; u++) J[u] = nbr[I * 27 + SLOTS * w + u];
#pragma unroll
        for (int u = 0; u < SLOTS; u++)
#pragma unroll
            for (int k = 0; k < K; k++) m[u][k] = blocks[blk<K>(N, I, SLOTS * w + u, k)];
#pragma unroll
        for (int u = 0; u < SLOTS; u++) {
            J[u] = J[u] < 0 ? I : J[u];
            P[u] = zc ? parent[J[u]] : 0;
#pragma unroll
            for (int c = 0; c < D; c++) zj[u][c] = z[(size_t)D * J[u] + c];
        }
        if (zc) {
#pragma unroll
            for (int u = 0; u < SLOTS; u++)
#pragma unroll
                for (int c = 0; c < D; c++) zj[u][c] += zc[(size_t)D * P[u] + c];
        }
#pragma unroll
        for (int u = 0; u < SLOTS; u++) {
            if (K == 9) {
                creal t[3];
                mat3_vec(m[u], zj[u], t);
#pragma unroll
                for (int c = 0; c < 3; c++) o[c] += t[c];
            } else {
#pragma unroll
                for (int c = 0; c < D; c++) o[c] += m[u][0] * zj[u][c];
            }
        }
    }
#pragma unroll
    for (int c = 0; c < D; c++) part[w][lane * D + c] = o[c];
    __syncthreads();
    creal s = 0;
    if (live) {
#pragma unroll
        for (int q = 0; q < WARPS; q++) s += part[q][t];
        s = rv - s;
        if (mode == 1) out[(size_t)D * Ic + c] = s;
        else if (K == 9) resS[t] = s;
    }
    if (mode == 1) return;
    if (K == 9) __syncthreads();
    if (!live) return;
    creal u;
    if (K == 9) {
        u = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) u += bi[k] * resS[ln * D + k];
    } else
        u = bi[0] * s;
    out[(size_t)D * Ic + c] = zi + omega * u;
}
// Residual of a small level AND its restriction to the next one (with that level's first sweep) in one launch: a CTA takes four
// coarse cells, whose children are at most 32 consecutive cells of the fine level — one chunk of k_coarse_apply<K, D, 1>'s layout
// (warp = stencil slot, lane = cell) — computes their residuals r - A z into shared memory instead of global memory, and adds
// them up per parent. Same arithmetic and summation orders as k_coarse_apply<K, D, 1> (mode 1) followed by k_restrict_coarse;
// one launch less per level and visit, which is what the latency-bound levels are made of.
constexpr int FUSE_G = 4;
template <int K, int D>
__global__ void __launch_bounds__(27 * 32) k_residual_restrict(const creal* __restrict__ blocks, const int* __restrict__ nbr, const creal* __restrict__ r,
                                                              const creal* __restrict__ z, int N, const int* __restrict__ firstChild, const creal* __restrict__ binvC,
                                                              const double* __restrict__ omegaP, int Nc, creal* __restrict__ rc, creal* __restrict__ zc) { pdl_wait();
    __shared__ creal part[27][32 * D];
    __shared__ creal resid[32 * D];
    __shared__ creal sums[FUSE_G * D];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
    const int c0 = blockIdx.x * FUSE_G, c1 = min(c0 + FUSE_G, Nc);
    const int f0 = firstChild[c0], f1 = firstChild[c1];
    const int I = f0 + lane;
    const int ln = t / D, c = t - D * ln;
    const int Ic = f0 + ln;
    const bool live = t < 32 * D && Ic < f1;
    const creal rv = live ? r[(size_t)D * Ic + c] : (creal)0;
    creal o[D];
#pragma unroll
    for (int c2 = 0; c2 < D; c2++) o[c2] = 0;
    if (I < f1) {
        creal m[K], zj[D];
        int J = nbr[I * 27 + w];
#pragma unroll
        for (int k = 0; k < K; k++) m[k] = blocks[blk<K>(N, I, w, k)];
        J = J < 0 ? I : J;
#pragma unroll
        for (int c2 = 0; c2 < D; c2++) zj[c2] = z[(size_t)D * J + c2];
        if (K == 9) {
            creal t3[3];
            mat3_vec(m, zj, t3);
#pragma unroll
            for (int c2 = 0; c2 < 3; c2++) o[c2] += t3[c2];
        } else {
#pragma unroll
            for (int c2 = 0; c2 < D; c2++) o[c2] += m[0] * zj[c2];
        }
    }
#pragma unroll
    for (int c2 = 0; c2 < D; c2++) part[w][lane * D + c2] = o[c2];
    __syncthreads();
    if (live) {
        creal sres = 0;
#pragma unroll
        for (int q = 0; q < 27; q++) sres += part[q][t];
        resid[t] = rv - sres;
    }
    __syncthreads();
    const int ci = t / D, cc = t - D * ci;
    const int cell = c0 + ci;
    const bool mine = t < FUSE_G * D && cell < c1;
    creal a = 0;
    if (mine) {
        for (int Ich = firstChild[cell]; Ich < firstChild[cell + 1]; Ich++) a += resid[(Ich - f0) * D + cc];
        sums[t] = a;
    }
    __syncthreads();
    if (!mine) return;
    const creal omega = (creal)*omegaP;
    creal u;
    if (K == 9) u = binvC[(size_t)(3 * cc) * Nc + cell] * sums[ci * D] + binvC[(size_t)(3 * cc + 1) * Nc + cell] * sums[ci * D + 1] + binvC[(size_t)(3 * cc + 2) * Nc + cell] * sums[ci * D + 2];
    else u = binvC[cell] * a;
    rc[(size_t)D * cell + cc] = a;
    zc[(size_t)D * cell + cc] = omega * u;
}

// Restriction between coarse levels (children of a cell are contiguous) with the first sweep of the coarser level.
// One thread per coarse cell.
template <int K, int D>
__global__ void k_restrict_coarse(const int* __restrict__ firstChild, const creal* __restrict__ rFine, int Ncoarse, const creal* __restrict__ binv,
                                  const double* __restrict__ omegaP, creal* __restrict__ rc, creal* __restrict__ zc, int c0 = 0, int c1 = -1) { pdl_wait();
    int Ip = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (Ip >= (c1 < 0 ? Ncoarse : c1)) return;
    const creal omega = (creal)*omegaP;
    creal a[D], o[D];
#pragma unroll
    for (int c = 0; c < D; c++) a[c] = 0;
    for (int I = firstChild[Ip]; I < firstChild[Ip + 1]; I++)
#pragma unroll
        for (int c = 0; c < D; c++) a[c] += rFine[(size_t)D * I + c];
    apply_binv<K, D>(binv, Ncoarse, Ip, a, o);
#pragma unroll
    for (int c = 0; c < D; c++) rc[(size_t)D * Ip + c] = a[c], zc[(size_t)D * Ip + c] = omega * o[c];
}
// z = omega * Binv * r on a level (partitioned mesh: the first sweep of level 1, after the restriction was all-reduced)
template <int K, int D>
__global__ void k_level_presmooth(const creal* __restrict__ binv, const creal* __restrict__ r, const double* __restrict__ omegaP, int N, creal* __restrict__ z) { pdl_wait();
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= N) return;
    const creal omega = (creal)*omegaP;
    creal v[D], o[D];
#pragma unroll
    for (int c = 0; c < D; c++) v[c] = r[(size_t)D * I + c];
    apply_binv<K, D>(binv, N, I, v, o);
#pragma unroll
    for (int c = 0; c < D; c++) z[(size_t)D * I + c] = omega * o[c];
}
template <int D>
__global__ void k_prolong_coarse(const int* __restrict__ parent, const creal* __restrict__ zc, int N, creal* __restrict__ z, int c0 = 0, int c1 = -1) { pdl_wait();
    int i = D * c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D * (c1 < 0 ? N : c1)) return;
    z[i] += zc[(size_t)D * parent[i / D] + i % D];
}
// Coarsest level: restriction into shared memory (every CTA repeats it, it is tiny), then z = M r with the dense inverse
// M (n x n, row-major): one warp per row, lanes across the columns. r and z are flat [n * C] = [cells][D].
constexpr int DENSE_MAX = 6 * 2 * COARSEST_CELLS;
constexpr int DENSE_CTAS = 8;
template <int D>
__global__ void __launch_bounds__(B) k_dense_restrict_apply(const int* __restrict__ firstChild, const creal* __restrict__ rFine, const creal* __restrict__ m, int n,
                                                           creal* __restrict__ z) { pdl_wait();
    constexpr int C = D == 3 ? 1 : 6;  // FLOW: n = 3 * cells, one right-hand side; SCALAR: n = cells, six
    __shared__ creal r[DENSE_MAX];
    const int total = n * C;  // = D * cells
    for (int i = threadIdx.x; i < total; i += B) {
        int Ip = i / D, c = i - D * Ip;
        creal a = 0;
        if (firstChild)
            for (int I = firstChild[Ip]; I < firstChild[Ip + 1]; I++) a += rFine[(size_t)D * I + c];
        else a = rFine[i];  // single-level hierarchy: already restricted
        r[i] = a;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * (B / 32) + (threadIdx.x >> 5); row < n; row += gridDim.x * (B / 32)) {
        creal acc[C];
#pragma unroll
        for (int c = 0; c < C; c++) acc[c] = 0;
        for (int k = lane; k < n; k += 32) {
            const creal mv = m[(size_t)row * n + k];
#pragma unroll
            for (int c = 0; c < C; c++) acc[c] += mv * r[k * C + c];
        }
#pragma unroll
        for (int c = 0; c < C; c++)
            for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        if (lane == 0)
#pragma unroll
            for (int c = 0; c < C; c++) z[row * C + c] = acc[c];
    }
}

// ----------------------------------------------------------- the small levels as ONE kernel on ONE cluster
//
// Below a few thousand cells a level is latency-bound: each of its kernels runs 4-9 us whatever it computes (launch ramp, then
// two or three DEPENDENT trips to global memory at ~1 000 cycles each: neighbour table -> neighbour's value -> ...), and a
// cycle visits the small levels of a hierarchy 30-50 times (W shape on the large levels). k_coarse_tail runs a whole visit of
// such a level — restriction from the level above, residual, the recursive visit of everything below down to the dense
// coarsest solve, W passes, post-smoothing — as one launch of ONE thread-block cluster (16 CTAs of 27 warps, one GPC):
//  * every level's cells are dealt to the CTAs in contiguous ranges; its vectors (r, iterate, scratch), neighbour table,
//    block inverses, parent / child links live in the owning CTA's SHARED memory for the whole launch, and so do the stencil
//    coefficients of the levels of <= 1 536 cells ("resident"; a larger level streams its coefficients from global memory
//    in every application — loads that depend on nothing and are issued ahead of everything else);
//  * a neighbour's value is read from the owner's shared memory (distributed shared memory, ~200 cycles), and the operations
//    of the visit are separated by the cluster's hardware barrier instead of a kernel boundary: no operation has a trip to
//    global memory on its critical path.
// Measured on the way here (profiles/r2c_tail_trace.txt): the same interpreter as a cooperative grid over all 148 SMs with a
// barrier through global memory costs 3.3-5 us per operation, and as one cluster with the vectors left in global memory
// 2.3-5 us — no better than a launch; the dependent global loads (and the fences that publish global stores) are the cost, not
// the launch.
// The host walks the same recursion as coarse_cycle() and writes the operations into the kernel's parameters (TailArgs::op);
// the kernel is an interpreter over them. The arithmetic of every operation, and the order of every sum, are those of the
// stand-alone kernels above (k_coarse_apply<K, D, 1>, k_restrict_coarse, k_prolong_coarse, k_dense_restrict_apply), so both
// paths give the same bits.
constexpr int TAIL_T = 27 * 32;
constexpr int TAIL_MAX_OPS = 200;
constexpr int TAIL_BATCH = 96;         // cells of a CTA's range per pass of an operator application
constexpr int TAIL_MAX_CTAS = 16;
enum : unsigned char { T_RESIDUAL = 0, T_RESTRICT = 1, T_DENSE = 2, T_PROLONG = 3, T_SMOOTH = 4, T_LOAD = 5 };
struct TailLevel {
    // global memory
    const creal* blocks;
    const int* nbr;
    const creal* binv;
    const int* parent;      // of this level's cells, on the next coarser level
    const int* firstChild;  // of this level's cells, on the next finer level
    creal* r;               // level vectors: read / written only for the level above the tail (its residual) and the tail's
    creal* z;               // first level (what comes in when there is no level above; the result)
    creal* t;
    int N, cpc;             // cells; cells per CTA (CTA k owns cells [k * cpc, (k + 1) * cpc))
    int resident;           // stencil coefficients staged in shared memory
    // word offsets of this level's pieces in every CTA's dynamic shared memory
    int oBlocks, oNbr, oBinv, oParent, oChild, oR, oBuf[2];
};
struct TailArgs {
    TailLevel lev[MAXL + 1];
    const creal* cinv;      // dense inverse of the coarsest level, nDense x nDense
    const double* omega;    // Multigrid::domega
    unsigned long long* trace;  // MOF_MG_TAIL_TRACE=1: %globaltimer of CTA 0 at the start of every operation and at the end
    int first, last;        // stencil levels first .. last - 1 and the dense level `last` are held in shared memory
    int oCinv, oPart, oRes, oDense;  // more word offsets: this CTA's rows of cinv, the slot products, residuals, the dense right-hand side
    int smemBytes;
    int nOps, nDense;
    unsigned char op[TAIL_MAX_OPS], lvl[TAIL_MAX_OPS], flip[TAIL_MAX_OPS];  // flip: bit 0 = where level lvl's iterate is, bit 1 = level lvl+1's
};

// Barrier over the whole cluster (= the whole grid) with release / acquire semantics: what a CTA wrote (to its shared memory,
// to global memory) before it is visible to every CTA after it.
__device__ __forceinline__ void tail_barrier() {
#ifdef MOF_HOST_EMULATION
    mof_emul::grid_sync();
#else
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
// The dynamic shared memory of CTA `rank` of the cluster, as a generic pointer.
__device__ __forceinline__ creal* tail_peer(creal* mine, int rank) {
#ifdef MOF_HOST_EMULATION
    return (creal*)mof_emul::peer_smem(mine, rank);
#else
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)mine), "r"(rank));
    return (creal*)out;
#endif
}
// (owner CTA, index in its range) of cell I of a level dealt in ranges of cpc cells, packed.
__device__ __forceinline__ int tail_pack(int I, int cpc) {
    const int owner = I / cpc;
    return (owner << 16) | (I - owner * cpc);
}

template <int K, int D>
__global__ void __launch_bounds__(TAIL_T, 1) k_coarse_tail(const TailArgs a) { pdl_wait();
    constexpr int C = D == 3 ? 1 : 6;
    constexpr int PROW = TAIL_BATCH * D;  // part[slot][cell of the batch][component]
#ifdef MOF_HOST_EMULATION
    creal* smem = (creal*)mof_emul::dynamic_smem((size_t)a.smemBytes);
#else
    extern __shared__ creal smem[];
#endif
    __shared__ creal* peers[TAIL_MAX_CTAS];
    const int rank = blockIdx.x, tid = threadIdx.x;
#ifdef MOF_HOST_EMULATION
    mof_emul::grid_sync();  // (every emulated CTA has registered its buffer)
#else
    if (a.trace && rank == 0 && tid == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        a.trace[255] = now;
    }
#endif
    if (tid < (int)gridDim.x) peers[tid] = tail_peer(smem, tid);
    creal* part = smem + a.oPart;
    creal* resS = smem + a.oRes;
    creal* dr = smem + a.oDense;
    int* ismem = reinterpret_cast<int*>(smem);

    // ---- staging: everything static of my cell ranges, once per launch. Eight independent loads per thread in flight at a time.
#define MOF_STAGE(total, SRC, DSTSTMT)                                   \
    for (int k0 = tid; k0 < (total); k0 += 8 * TAIL_T) {                 \
        creal v8[8];                                                     \
        _Pragma("unroll") for (int u8 = 0; u8 < 8; u8++) {               \
            const int k = k0 + u8 * TAIL_T;                              \
            if (k < (total)) v8[u8] = (SRC);                             \
        }                                                                \
        _Pragma("unroll") for (int u8 = 0; u8 < 8; u8++) {               \
            const int k = k0 + u8 * TAIL_T;                              \
            const creal v = v8[u8];                                      \
            if (k < (total)) { DSTSTMT; }                                \
        }                                                                \
    }
    for (int l = a.first; l <= a.last; l++) {
        const TailLevel& L = a.lev[l];
        const int c0 = rank * L.cpc, n = max(0, min(L.N, c0 + L.cpc) - c0);
        if (l == a.last) {  // the dense level: the child links of ALL its cells (every CTA restricts the whole right-hand side), my rows of the inverse
            for (int i = tid; i <= L.N; i += TAIL_T) ismem[L.oChild + i] = L.firstChild[i];
            const int bs = D == 3 ? 3 : 1, rows = bs * n, nd = a.nDense;
            MOF_STAGE(rows * nd, a.cinv[(size_t)bs * c0 * nd + k], smem[a.oCinv + k] = v)
            continue;
        }
        if (n <= 0) continue;
        if (L.firstChild)
            for (int i = tid; i <= n; i += TAIL_T) ismem[L.oChild + i] = L.firstChild[c0 + i];
        const int cpcUp = a.lev[l + 1].cpc;
        for (int k = tid; k < 27 * n; k += TAIL_T) {
            const int i = k / 27, s = k - 27 * i;
            int J = L.nbr[(size_t)c0 * 27 + k];
            ismem[L.oNbr + s * L.cpc + i] = tail_pack(J < 0 ? c0 + i : J, L.cpc);
        }
        MOF_STAGE(K * n, L.binv[(size_t)(k / n) * L.N + c0 + k % n], smem[L.oBinv + (k / n) * L.cpc + k % n] = v)
        for (int i = tid; i < n; i += TAIL_T) ismem[L.oParent + i] = tail_pack(L.parent[c0 + i], cpcUp);
        if (L.resident) MOF_STAGE(27 * K * n, L.blocks[(size_t)(k / n) * L.N + c0 + k % n], smem[L.oBlocks + (k / n) * L.cpc + k % n] = v)
    }
#undef MOF_STAGE
    __syncthreads();

    const int w = tid >> 5, lane = tid & 31;
    for (int op = 0; op <= a.nOps; op++) {
        if (op) tail_barrier();
#ifndef MOF_HOST_EMULATION
        if (a.trace && rank == 0 && tid == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            a.trace[op] = now;
        }
#endif
        if (op == a.nOps) break;
        const int l = a.lvl[op], fl = a.flip[op] & 1, fu = (a.flip[op] >> 1) & 1;
        const TailLevel& L = a.lev[l];
        const TailLevel& U = a.lev[l + 1];
        const int what = a.op[op];
        if (what == T_RESIDUAL || what == T_SMOOTH) {
            // mode 1: scratch = r - A z ; mode 2: scratch = z + omega Binv (r - A z). Warp = stencil slot, lane = cell; the 27 products of
            // a cell meet in shared memory and are added in slot order.
            const int c0 = rank * L.cpc, n = max(0, min(L.N, c0 + L.cpc) - c0);
            const creal omega = (creal)a.omega[1 + l];
            const int oZ = L.oBuf[fl];
            creal* out = smem + L.oBuf[1 - fl];
            const bool corr = what == T_SMOOTH;  // post-smoothing adds the coarse correction on the fly: z stands for z + P zc
            const int oZc = U.oBuf[fu];
            for (int b0 = 0; b0 < n; b0 += TAIL_BATCH) {
                const int nb = min(TAIL_BATCH, n - b0);
                for (int i = lane; i < nb; i += 32) {
                    creal m[K], zj[D];
                    if (L.resident) {
#pragma unroll
                        for (int k = 0; k < K; k++) m[k] = smem[L.oBlocks + (w * K + k) * L.cpc + b0 + i];
                    } else {
#pragma unroll
                        for (int k = 0; k < K; k++) m[k] = L.blocks[(size_t)(w * K + k) * L.N + c0 + b0 + i];
                    }
                    const int pj = ismem[L.oNbr + w * L.cpc + b0 + i];
                    const creal* base = peers[pj >> 16];
                    const creal* zsrc = base + oZ + (pj & 0xffff) * D;
#pragma unroll
                    for (int c = 0; c < D; c++) zj[c] = zsrc[c];
                    if (corr) {
                        const int pp = reinterpret_cast<const int*>(base)[L.oParent + (pj & 0xffff)];
                        const creal* csrc = peers[pp >> 16] + oZc + (pp & 0xffff) * D;
#pragma unroll
                        for (int c = 0; c < D; c++) zj[c] += csrc[c];
                    }
                    creal* dst = part + w * PROW + i * D;
                    if (K == 9) {
                        creal t3[3];
                        mat3_vec(m, zj, t3);
#pragma unroll
                        for (int c = 0; c < 3; c++) dst[c] = t3[c];
                    } else {
#pragma unroll
                        for (int c = 0; c < D; c++) dst[c] = m[0] * zj[c];
                    }
                }
                __syncthreads();
                creal sres = 0;
                const bool live = tid < nb * D;
                const int i = tid / D, c = tid - D * i;
                if (live) {
#pragma unroll
                    for (int q = 0; q < 27; q++) sres += part[q * PROW + tid];
                    sres = smem[L.oR + (b0 + i) * D + c] - sres;
                    if (what == T_RESIDUAL) out[(b0 + i) * D + c] = sres;
                    else if (K == 9) resS[tid] = sres;
                }
                if (what == T_SMOOTH) {
                    if (K == 9) __syncthreads();
                    if (live) {
                        creal u;
                        if (K == 9) {
                            u = 0;
#pragma unroll
                            for (int k = 0; k < 3; k++) u += smem[L.oBinv + (3 * c + k) * L.cpc + b0 + i] * resS[i * D + k];
                        } else
                            u = smem[L.oBinv + b0 + i] * sres;
                        const int pp = ismem[L.oParent + b0 + i];
                        const creal zi = smem[oZ + (b0 + i) * D + c] + peers[pp >> 16][oZc + (pp & 0xffff) * D + c];
                        out[(b0 + i) * D + c] = zi + omega * u;
                    }
                }
                __syncthreads();  // part / resS are written again by the next batch
            }
        } else if (what == T_RESTRICT) {
            // the level below's residual (sum over the children, which are consecutive cells) and first sweep; the children's
            // residuals come from the level above the tail in global memory (l < first) or from the owners' shared memory
            const int c0 = rank * U.cpc, n = max(0, min(U.N, c0 + U.cpc) - c0);
            const creal omega = (creal)a.omega[2 + l];
            const bool fromGlobal = l < a.first;
            const int oT = fromGlobal ? 0 : L.oBuf[1 - fl];
            for (int i = tid; i < n; i += TAIL_T) {
                creal sum[D], o[D];
#pragma unroll
                for (int c = 0; c < D; c++) sum[c] = 0;
                const int f0 = ismem[U.oChild + i], f1 = ismem[U.oChild + i + 1];
                if (fromGlobal) {
                    for (int I = f0; I < f1; I++)
#pragma unroll
                        for (int c = 0; c < D; c++) sum[c] += __ldcg(L.t + (size_t)D * I + c);
                } else {
                    int owner = f0 / L.cpc, local = f0 - owner * L.cpc;
                    for (int I = f0; I < f1; I++) {
                        const creal* src = peers[owner] + oT + local * D;
#pragma unroll
                        for (int c = 0; c < D; c++) sum[c] += src[c];
                        if (++local == L.cpc) local = 0, owner++;
                    }
                }
                if (K == 9) {
                    creal m[9];
#pragma unroll
                    for (int k = 0; k < 9; k++) m[k] = smem[U.oBinv + k * U.cpc + i];
                    mat3_vec(m, sum, o);
                } else {
                    const creal wv = smem[U.oBinv + i];
#pragma unroll
                    for (int c = 0; c < D; c++) o[c] = wv * sum[c];
                }
#pragma unroll
                for (int c = 0; c < D; c++) smem[U.oR + i * D + c] = sum[c], smem[U.oBuf[fu] + i * D + c] = omega * o[c];
            }
        } else if (what == T_DENSE) {
            // coarsest level: every CTA restricts the whole right-hand side into its shared memory (it is tiny), then applies its rows of
            // the dense inverse: one warp per row, lanes across the columns
            const int nd = a.nDense, total = nd * C;
            const int c0 = rank * U.cpc, n = max(0, min(U.N, c0 + U.cpc) - c0);
            const int bs = D == 3 ? 3 : 1;
            const bool fromGlobal = l < a.first;
            const int oT = fromGlobal ? 0 : L.oBuf[1 - fl];
            if (n > 0) {
                for (int k = tid; k < total; k += TAIL_T) {
                    const int Ip = k / D, c = k - D * Ip;
                    const int f0 = ismem[U.oChild + Ip], f1 = ismem[U.oChild + Ip + 1];
                    creal sum = 0;
                    if (fromGlobal) {
                        for (int I = f0; I < f1; I++) sum += __ldcg(L.t + (size_t)D * I + c);
                    } else {
                        int owner = f0 / L.cpc, local = f0 - owner * L.cpc;
                        for (int I = f0; I < f1; I++) {
                            sum += peers[owner][oT + local * D + c];
                            if (++local == L.cpc) local = 0, owner++;
                        }
                    }
                    dr[k] = sum;
                }
                __syncthreads();
                for (int row = w; row < bs * n; row += TAIL_T / 32) {
                    creal acc[C];
#pragma unroll
                    for (int c = 0; c < C; c++) acc[c] = 0;
                    for (int k = lane; k < nd; k += 32) {
                        const creal mv = smem[a.oCinv + row * nd + k];
#pragma unroll
                        for (int c = 0; c < C; c++) acc[c] += mv * dr[k * C + c];
                    }
#pragma unroll
                    for (int c = 0; c < C; c++)
                        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
                    if (lane == 0)
#pragma unroll
                        for (int c = 0; c < C; c++) smem[U.oBuf[fu] + row * C + c] = acc[c];
                }
            }
        } else if (what == T_PROLONG) {  // z += P zc (before post-smoothing, and between the passes of a W visit)
            const int c0 = rank * L.cpc, n = max(0, min(L.N, c0 + L.cpc) - c0);
            for (int k = tid; k < n * D; k += TAIL_T) {
                const int i = k / D, c = k - D * i;
                const int pp = ismem[L.oParent + i];
                smem[L.oBuf[fl] + k] += peers[pp >> 16][U.oBuf[fu] + (pp & 0xffff) * D + c];
            }
        } else {  // T_LOAD: the tail starts at the hierarchy's first level — its residual and first sweep are in global memory
            const int c0 = rank * L.cpc, n = max(0, min(L.N, c0 + L.cpc) - c0);
            for (int k = tid; k < n * D; k += TAIL_T) {
                smem[L.oR + k] = __ldcg(L.r + (size_t)D * c0 + k);
                smem[L.oBuf[fl] + k] = __ldcg(L.z + (size_t)D * c0 + k);
            }
        }
    }
    // the result: the first level's iterate, for the level above (or the fine level's prolongation)
    {
        const TailLevel& L = a.lev[a.first];
        const int fin = a.flip[a.nOps] & 1;
        const int c0 = rank * L.cpc, n = max(0, min(L.N, c0 + L.cpc) - c0);
        for (int k = tid; k < n * D; k += TAIL_T) L.z[(size_t)D * c0 + k] = smem[L.oBuf[fin] + k];
    }
}

// ------------------------------------------------------------------------------------- PCG kernels

template <class TA, class TB>
__global__ void k_dot_partial(const TA* __restrict__ a, const TB* __restrict__ b, long long n, double* __restrict__ partial) { pdl_wait();
    __shared__ double sh[B];
    double s = 0;
    for (long long i = (long long)blockIdx.x * B + threadIdx.x; i < n; i += (long long)gridDim.x * B) s += (double)a[i] * (double)b[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// Folds the partials (fixed order) into scal[slot] and derives the PCG scalar that depends on it.
__global__ void k_fold(const double* __restrict__ partial, int np, int slot, double* __restrict__ scal) { pdl_wait(); fold_partials(partial, np, slot, scal); }
// x += alpha p ; r -= alpha q ; partial(r.r) ; and the cycle's pre-smoothing of the new residual, z = omega * dinv * r
// (the last CTA folds r.r into scal[S_RR])
__global__ void k_update_xr(const double* __restrict__ p, const double* __restrict__ q, long long n, double* __restrict__ x, double* __restrict__ r,
                            const creal* __restrict__ dinv, const double* __restrict__ omegaP, int nrhs, creal* __restrict__ z, Fold f) { pdl_wait();
    const double alpha = f.scal[S_ALPHA], omega = *omegaP;
    double s = 0;
    const long long stride = (long long)gridDim.x * B;
    // two elements per trip, all ten loads issued before the first store (the sum keeps the order of the one-element loop)
    for (long long i = (long long)blockIdx.x * B + threadIdx.x; i < n; i += 2 * stride) {
        const long long j = i + stride;
        const bool two = j < n;
        const double r0 = r[i], q0 = q[i], x0 = x[i], p0 = p[i], d0 = (double)dinv[i / nrhs];
        double r1 = 0, q1 = 0, x1 = 0, p1 = 0, d1 = 0;
        if (two) r1 = r[j], q1 = q[j], x1 = x[j], p1 = p[j], d1 = (double)dinv[j / nrhs];
        const double rv0 = r0 - alpha * q0;
        x[i] = x0 + alpha * p0;
        r[i] = rv0;
        z[i] = (creal)(omega * d0 * rv0);
        s += rv0 * rv0;
        if (two) {
            const double rv1 = r1 - alpha * q1;
            x[j] = x1 + alpha * p1;
            r[j] = rv1;
            z[j] = (creal)(omega * d1 * rv1);
            s += rv1 * rv1;
        }
    }
    cta_partial(s, f.partial, f.slot, f.scal, f.counter);
}
__global__ void k_direction(const creal* __restrict__ z, const double* __restrict__ scal, long long n, int first, double* __restrict__ p) { pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = first ? (double)z[i] : (double)z[i] + scal[S_BETA] * p[i];
}
// Loop control of a captured solve, one thread each.
// Before a solve: the threshold from b.b (already in scal), the counters.
__global__ void k_pcg_begin(double* __restrict__ scal, double tol2, int maxIters) { pdl_wait();
    scal[S_THR] = tol2 * scal[S_BB], scal[S_IT] = 0, scal[S_MAXIT] = maxIters, scal[S_RR0] = 0;
}
// In front of the loop (enter = 1): run it at all? At the end of a pass of two iterations (enter = 0): another one? The
// same test the host made between graph replays: both residual norms of the pass above the threshold, finite, iterations left.
__global__ void k_pcg_continue(cudaGraphConditionalHandle loop, double* __restrict__ scal, int enter) { pdl_wait();
    const double thr = scal[S_THR], rr = scal[S_RR];
    bool go = scal[S_BB] > 0 && rr > thr && rr < 1e300;  // (also false for NaN)
    if (!enter) {
        scal[S_IT] += 2;
        go = go && scal[S_RR0] > thr;
    }
    go = go && scal[S_IT] < scal[S_MAXIT];
    cudaGraphSetConditional(loop, go ? 1u : 0u);
}
// Power iteration support (spectral radius of Minv A per level, for the Jacobi damping).
__global__ void k_pseudo_random(long long n, creal* __restrict__ v) { pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = (unsigned)i * 2654435761u + 12345u;
    h ^= h >> 15, h *= 2246822519u, h ^= h >> 13;
    v[i] = (creal)((double)(h & 0xffff) / 32768. - 1.);
}
__global__ void k_normalise(const creal* __restrict__ t, const double* __restrict__ scal, int slot, long long n, creal* __restrict__ v) { pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = scal[slot];
    v[i] = (creal)(s > 0 ? (double)t[i] / sqrt(s) : 0.);
}

void drop_graphs(Multigrid& mg) {
    for (PcgGraph& g : mg.graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    mg.graphs.clear();
}

void release_mg(Multigrid* mg) {
    if (!mg) return;
    drop_graphs(*mg);
    mg->tailTrace.release();
    for (MgLevel& l : mg->lev) {
        l.code.release(), l.nbr.release(), l.parent.release(), l.firstChild.release(), l.blocks.release(), l.cblocks.release(), l.binv.release(), l.r.release(),
            l.z.release(), l.t.release();
    }
    mg->evec.release(), mg->cevec.release(), mg->cevecAgg.release(), mg->agg.release(), mg->aggPtr.release(), mg->aggList.release(), mg->slotOf.release(), mg->cinv.release();
    mg->fval.release(), mg->fdinv.release(), mg->fvalSell.release();
    mg->fz.release(), mg->fz2.release(), mg->ft.release(), mg->fr.release(), mg->fp.release(), mg->fq.release(), mg->partial.release(), mg->scal.release(), mg->counter.release(), mg->domega.release(), mg->eig.release(), mg->chebD.release(), mg->chebR.release(), mg->chebC.release();
    delete mg;
}

// ------------------------------------------------------------------------------------------ setup

// Octree over `n` points (device, [n][3]) whose coupled pairs are at most maxEdge apart: levels, neighbour tables,
// parent/child links, aggregate of every point and sorted member lists. Leaves mg.K == 0 if the scheme does not apply.
int build_octree(mof_ctx* ctx, Multigrid& mg, const double* pts, int n, double maxEdge, int target) {
    const int D = mg.dofs(), Kc = mg.comps();
    for (int k = 0; k < 3; k++) {
        MOF_TRY(minmax(ctx, mg, pts, n, 3, k, 0, mg.scal.p + k));
        MOF_TRY(minmax(ctx, mg, pts, n, 3, k, 1, mg.scal.p + 3 + k));
    }
    double h[6];
    MOF_CUDA(cudaMemcpyAsync(h, mg.scal.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    double ext = 0;
    for (int k = 0; k < 3; k++) ext = std::max(ext, h[3 + k] - h[k]);
    if (!(ext > 0) || !(maxEdge > 0)) return MOF_OK;
    ext *= 1.0001;
    GridMap gm;
    for (int k = 0; k < 3; k++) gm.lo[k] = h[k];
    gm.inv = 1. / ext;
    // finest admissible level: cells at least as wide as the longest edge
    int Lmax = 1;
    while (Lmax < MAXL && ext / (double)(1 << (Lmax + 1)) >= maxEdge) Lmax++;
    if (Lmax < 2) return MOF_OK;
    DBuf<int> occ[MAXL + 1], rank[MAXL + 1];
    auto freeAll = [&]() { for (int L = 0; L <= MAXL; L++) occ[L].release(), rank[L].release(); };
    auto levelRanks = [&](int L, int from) -> int {  // occupancy + Morton rank of level L, from the points (from < 0) or from level `from`
        long long cells = 1ll << (3 * L);
        MOF_CUDA(occ[L].alloc(cells + 1));
        MOF_CUDA(rank[L].alloc(cells + 1));
        MOF_CUDA(cudaMemsetAsync(occ[L].p, 0, sizeof(int) * (cells + 1), ctx->stream));
        if (from < 0) MOF_LAUNCH(k_mark_cells, blocks_for(n, B), B, 0, gm, pts, n, L, occ[L].p);
        else MOF_LAUNCH(k_mark_parents, blocks_for(1ll << (3 * from), B), B, 0, occ[from].p, 1ll << (3 * from), occ[L].p);
        return exclusive_scan_int(ctx, occ[L].p, rank[L].p, (int)cells + 1, nullptr);
    };
    // occupancy of a surface grows ~4x per level: probe one level, extrapolate to ~target unknowns per cell
    const int probe = std::min(Lmax, 5);
    int rc = levelRanks(probe, -1);
    if (rc != MOF_OK) { freeAll(); return rc; }
    int nProbe = 0;
    MOF_CUDA(read_back(ctx, &nProbe, rank[probe].p + (1ll << (3 * probe))));
    occ[probe].release(), rank[probe].release();
    int L1 = probe + (int)std::lround(std::log((double)n / ((double)target * std::max(nProbe, 1))) / std::log(4.0));
    L1 = std::max(2, std::min(Lmax, L1));
    for (int L = L1; L >= 1; L--) {
        rc = levelRanks(L, L == L1 ? -1 : L + 1);
        if (rc != MOF_OK) { freeAll(); return rc; }
    }
    std::vector<int> counts(L1 + 1, 0);
    for (int L = L1; L >= 1; L--) MOF_CUDA(read_back(ctx, &counts[L], rank[L].p + (1ll << (3 * L))));
    int Lc = L1;
    while (Lc > 1 && counts[Lc] > COARSEST_CELLS) Lc--;
    if (counts[Lc] > 2 * COARSEST_CELLS) { freeAll(); return MOF_OK; }
    // The grid levels of the hierarchy. MOF_MG_SKIP_CELLS (default 0: none): below that many cells every other octree level is left
    // out (cells coarsen 64-fold in volume, ~16-fold in number on a surface) — the small levels cost a fixed 4-9 us per kernel whatever
    // their size, and a cycle visits them 4-16 times. The cycle's shape (W levels) is that of the full hierarchy.
    const int skipCells = env_int(mg.kind == MG_FLOW ? "MOF_MG_SKIP_CELLS" : "MOF_MG_SKIP_CELLS_SCALAR", 0);
    std::vector<int> gridLevels(1, L1);
    for (int L = L1; L > Lc;) {
        L -= (counts[L] <= skipCells && L - 2 >= Lc) ? 2 : 1;
        gridLevels.push_back(L);
    }
    mg.fullDepth = L1 - Lc + 1;
    const int K = (int)gridLevels.size();
    for (size_t l = K; l < mg.lev.size(); l++) {  // a previous, deeper hierarchy: give the extra levels back
        MgLevel& o = mg.lev[l];
        o.code.release(), o.nbr.release(), o.parent.release(), o.firstChild.release(), o.blocks.release(), o.cblocks.release(), o.binv.release(), o.r.release(),
            o.z.release(), o.t.release();
    }
    mg.lev.resize(K);
    for (int l = 0; l < K; l++) {
        MgLevel& lv = mg.lev[l];
        int L = gridLevels[l];
        lv.gridLevel = L, lv.N = counts[L];
        long long cells = 1ll << (3 * L);
        MOF_CUDA(lv.code.alloc(lv.N));
        MOF_CUDA(lv.nbr.alloc(27ull * lv.N));
        MOF_CUDA(lv.blocks.alloc(27ull * Kc * lv.N));
        MOF_CUDA(lv.cblocks.alloc(27ull * Kc * lv.N));
        MOF_CUDA(lv.binv.alloc((size_t)Kc * lv.N));
        MOF_CUDA(lv.r.alloc((size_t)D * lv.N));
        MOF_CUDA(lv.z.alloc((size_t)D * lv.N));
        MOF_CUDA(lv.t.alloc((size_t)D * lv.N));
        MOF_LAUNCH(k_node_codes, blocks_for(cells, B), B, 0, occ[L].p, rank[L].p, cells, lv.code.p);
        MOF_LAUNCH(k_neighbours, blocks_for(27ll * lv.N, B), B, 0, lv.code.p, occ[L].p, rank[L].p, lv.N, L, lv.nbr.p);
    }
    for (int l = 0; l + 1 < K; l++) {
        MgLevel& lv = mg.lev[l];
        MgLevel& up = mg.lev[l + 1];
        MOF_CUDA(lv.parent.alloc(lv.N));
        MOF_CUDA(up.firstChild.alloc(up.N + 1));
        MOF_LAUNCH(k_parents, blocks_for(lv.N + 1, B), B, 0, lv.code.p, rank[up.gridLevel].p, lv.N, up.N, 3 * (lv.gridLevel - up.gridLevel), lv.parent.p, up.firstChild.p);
    }
    MgLevel& l1 = mg.lev[0];
    MOF_CUDA(mg.agg.alloc(n));
    MOF_CUDA(mg.aggPtr.alloc(l1.N + 1));
    MOF_CUDA(mg.aggList.alloc(n));
    DBuf<int> cnt, cursor, unsorted;
    MOF_CUDA(cnt.alloc(l1.N + 1));
    MOF_CUDA(cursor.alloc(l1.N + 1));
    MOF_CUDA(unsorted.alloc(n));
    MOF_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int) * (l1.N + 1), ctx->stream));
    MOF_LAUNCH(k_point_aggregate, blocks_for(n, B), B, 0, gm, pts, rank[L1].p, n, L1, mg.agg.p, cnt.p);
    rc = exclusive_scan_int(ctx, cnt.p, mg.aggPtr.p, l1.N + 1, nullptr);
    if (rc == MOF_OK) {
        cudaMemcpyAsync(cursor.p, mg.aggPtr.p, sizeof(int) * (l1.N + 1), cudaMemcpyDeviceToDevice, ctx->stream);
        MOF_LAUNCH(k_aggregate_fill, blocks_for(n, B), B, 0, mg.agg.p, n, cursor.p, unsorted.p);
        MOF_LAUNCH(k_aggregate_sort, blocks_for(32ll * l1.N, B), B, 0, mg.aggPtr.p, l1.N, unsorted.p, mg.aggList.p);
        cudaStreamSynchronize(ctx->stream);
    }
    cnt.release(), cursor.release(), unsorted.release();
    freeAll();
    if (rc != MOF_OK) return rc;
    mg.K = K;
    return MOF_OK;
}

int plan_tail(mof_ctx* ctx, Multigrid& mg);  // below, with the cycle

// Over-correction: piecewise-constant aggregation makes the Galerkin operator of a Laplacian-like system too stiff (the energy
// of a step function over-estimates that of the smooth function it stands for), so its coarse corrections come out too small.
// Dividing every coarse operator by a factor (relative to the Galerkin product of the level above) enlarges them; the cycle
// stays symmetric positive definite. MOF_MG_OVERCORRECT / MOF_MG_OVERCORRECT_SCALAR (default 1: plain Galerkin).
double coarse_scale(const Multigrid& mg, int level) {
    (void)level;
    static const double flow = [] { const char* e = getenv("MOF_MG_OVERCORRECT"); return e && *e ? atof(e) : 1.0; }();
    static const double scalar = [] { const char* e = getenv("MOF_MG_OVERCORRECT_SCALAR"); return e && *e ? atof(e) : 1.0; }();
    const double a = mg.kind == MG_FLOW ? flow : scalar;
    return a > 0 ? 1. / a : 1.;
}

int alloc_common(mof_ctx* ctx, Multigrid& mg) {
    const size_t len = mg.fineLen();
    MOF_CUDA(mg.fz.alloc(len));
    MOF_CUDA(mg.fz2.alloc(len));
    MOF_CUDA(mg.ft.alloc(len));
    MOF_CUDA(mg.fr.alloc(len));
    MOF_CUDA(mg.fp.alloc(len));
    MOF_CUDA(mg.fq.alloc(len));
    MOF_CUDA(mg.fval.alloc(mg.kind == MG_FLOW ? (size_t)ctx->wPadded : (size_t)ctx->nnzS));
    MOF_CUDA(mg.fdinv.alloc((size_t)mg.nFine));
    if (mg.kind == MG_SCALAR && ctx->sPadded > 0) MOF_CUDA(mg.fvalSell.alloc((size_t)ctx->sPadded));
    const int nc = (mg.kind == MG_FLOW ? 3 : 1) * mg.lev.back().N;
    MOF_CUDA(mg.cinv.alloc((size_t)nc * nc));
    if (mg.kind == MG_FLOW) {
        MOF_CUDA(mg.cevec.alloc(3ull * mg.nFine));
        MOF_LAUNCH(k_to_creal, kSMs * 8, B, 0, mg.evec.p, 3ll * mg.nFine, mg.cevec.p);
        MOF_CUDA(mg.cevecAgg.alloc(3ull * mg.nFine));
        MOF_LAUNCH(k_gather_evec, blocks_for(3ll * mg.nFine, B), B, 0, (const int*)mg.aggList.p, (const creal*)mg.cevec.p, mg.nFine, mg.cevecAgg.p);
    }
    return MOF_OK;
}

int upload_omegas(mof_ctx* ctx, Multigrid& mg) {
    double* h = mg.hostOmega;
    for (int i = 0; i < OM_COUNT; i++) h[i] = 0.6;
    h[0] = mg.omega0;
    for (int l = 0; l < mg.K && 1 + l < OM_MINUS_ONE; l++) h[1 + l] = mg.lev[l].omega;
    h[OM_MINUS_ONE] = -1., h[OM_ZERO] = 0.;
    MOF_CUDA(cudaMemcpyAsync(mg.domega.p, h, sizeof(double) * OM_COUNT, cudaMemcpyHostToDevice, ctx->stream));
    return MOF_OK;
}

// A hierarchy object is kept from mesh to mesh: its buffers are re-sized in place (DBuf keeps memory that still fits).
Multigrid* new_mg(mof_ctx* ctx, Multigrid* old, MgKind kind, int nFine, int* rcOut) {
    Multigrid* mg = old ? old : new Multigrid();
    mg->usable = false, mg->K = 0;
    drop_graphs(*mg);  // captured for the previous mesh's buffers
    mg->eigValid = false;
    mg->tailStart = -1;
    mg->kind = kind, mg->nFine = nFine, mg->nrhs = kind == MG_FLOW ? 1 : 6;
    mg->gamma = std::max(1, std::min(2, env_int("MOF_MG_GAMMA", 2)));
    {
        for (int& c : mg->coarseSweeps) c = 1;
        const char* e = getenv(kind == MG_FLOW ? "MOF_MG_COARSE_SWEEPS" : "MOF_MG_COARSE_SWEEPS_SCALAR");
        for (int l = 0; e && *e && l <= MAXL; l++) {
            mg->coarseSweeps[l] = std::max(1, std::min(4, atoi(e)));
            e = strchr(e, ',');
            if (e) e++;
        }
    }
    mg->fineSweeps = std::max(1, std::min(4, env_int(kind == MG_FLOW ? "MOF_MG_FINE_SWEEPS" : "MOF_MG_FINE_SWEEPS_SCALAR", 1)));
    mg->gammaLevels = env_int(kind == MG_FLOW ? "MOF_MG_GAMMA_LEVELS" : "MOF_MG_GAMMA_LEVELS_SCALAR", -1);  // < 0: by depth, see coarse_cycle
    cudaError_t e = mg->partial.alloc(8192);  // per-CTA partials: NBLK of ours, or the persistent-grid size of k_spmv_dot
    if (e == cudaSuccess) e = mg->scal.alloc(64);
    if (e == cudaSuccess) e = mg->counter.alloc(4);
    if (e == cudaSuccess) e = mg->domega.alloc(OM_COUNT);
    if (e == cudaSuccess && upload_omegas(ctx, *mg) != MOF_OK) e = cudaErrorInvalidValue;
    if (e == cudaSuccess) e = cudaMemsetAsync(mg->counter.p, 0, 4 * sizeof(unsigned), ctx->stream);
    if (e == cudaSuccess && env_int("MOF_MG_TAIL_TRACE", 0)) e = mg->tailTrace.alloc(256);
    mg->hostRR = ctx->pinned + (kind == MG_FLOW ? 0 : 64);
    *rcOut = e == cudaSuccess ? MOF_OK : cuda_fail(ctx, e, "multigrid workspace");
    return mg;
}

}  // namespace

void mg_destroy(mof_ctx* ctx) {
    release_mg(ctx->mg), release_mg(ctx->mgs);
    ctx->mg = ctx->mgs = nullptr;
}

void mg_new_pair(mof_ctx* ctx) {
    if (ctx->mg) ctx->mg->eigValid = false;
    if (ctx->mgs) ctx->mgs->eigValid = false;
}

// Mesh-dependent part of both hierarchies.
int mg_setup_mesh(mof_ctx* ctx) {
    PhaseTimer pt(ctx);
    const int E = ctx->E, V = ctx->V;
    int rc = MOF_OK;
    ctx->mg = new_mg(ctx, ctx->mg, MG_FLOW, E, &rc);
    if (rc != MOF_OK) return rc;
    ctx->mgs = new_mg(ctx, ctx->mgs, MG_SCALAR, V, &rc);
    if (rc != MOF_OK) return rc;
    Multigrid& mf = *ctx->mg;
    Multigrid& ms = *ctx->mgs;
    // edge geometry: vectors (FLOW prolongation), midpoints (FLOW octree), longest edge (both)
    DBuf<double>& emid = ctx->dtmp2;
    MOF_CUDA(mf.evec.alloc(3ull * E));
    MOF_CUDA(emid.reserve(3ull * E));
    MOF_CUDA(ctx->dtmp0.reserve((size_t)E));
    MOF_LAUNCH(k_edge_geometry, blocks_for(E, B), B, 0, ctx->pos.p, ctx->tri.p, ctx->expanded.p, E, mf.evec.p, emid.p, ctx->dtmp0.p);
    MOF_TRY(minmax(ctx, mf, ctx->dtmp0.p, E, 1, 0, 1, mf.scal.p + 6));
    double maxLen2 = 0;
    MOF_CUDA(cudaMemcpyAsync(&maxLen2, mf.scal.p + 6, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    const double maxEdge = std::sqrt(maxLen2);

    if (env_int("MOF_FLOW_MG", 1)) {
        rc = build_octree(ctx, mf, emid.p, E, maxEdge, env_int("MOF_MG_TARGET", 40));
        if (rc == MOF_OK && mf.K > 0) {
            MOF_CUDA(mf.slotOf.alloc((size_t)ctx->wPadded));
            MOF_CUDA(cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * 16, ctx->stream));
            MOF_LAUNCH(k_entry_slots_flow, blocks_for(32ll * ctx->wSlices, B), B, 0, ctx->wRowptr.p, ctx->wSliceBase.p, ctx->wCol.p, mf.agg.p, mf.lev[0].code.p, E,
                       ctx->wSlices, mf.slotOf.p, ctx->flags.p);
            int hflag = 0;
            MOF_CUDA(cudaMemcpyAsync(&hflag, ctx->flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            MOF_CUDA(cudaStreamSynchronize(ctx->stream));
            if (!hflag) {
                rc = alloc_common(ctx, mf);
                if (rc == MOF_OK) rc = plan_tail(ctx, mf);
                mf.usable = rc == MOF_OK;
            }
        }
    }
    if (rc != MOF_OK) return rc;
    pt.mark("  flow hierarchy");

    if (env_int("MOF_SCALAR_MG", 1)) {
        rc = build_octree(ctx, ms, ctx->pos.p, V, maxEdge, env_int("MOF_MG_TARGET_SCALAR", 14));
        if (rc == MOF_OK && ms.K > 0) {
            MOF_CUDA(ms.slotOf.alloc((size_t)ctx->nnzS));
            MOF_CUDA(cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * 16, ctx->stream));
            MOF_LAUNCH(k_entry_slots_scalar, blocks_for(V, B), B, 0, ctx->sRowptr.p, ctx->sCol.p, ms.agg.p, ms.lev[0].code.p, V, ms.slotOf.p, ctx->flags.p);
            int hflag = 0;
            MOF_CUDA(cudaMemcpyAsync(&hflag, ctx->flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            MOF_CUDA(cudaStreamSynchronize(ctx->stream));
            if (!hflag) {
                rc = alloc_common(ctx, ms);
                if (rc == MOF_OK) rc = plan_tail(ctx, ms);
                ms.usable = rc == MOF_OK;
            }
        }
    }
    pt.mark("  scalar hierarchy");
    return rc;
}

bool mg_flow_usable(const mof_ctx* ctx) { return ctx->mg && ctx->mg->usable; }
bool mg_scalar_usable(const mof_ctx* ctx) { return ctx->mgs && ctx->mgs->usable; }

namespace {

constexpr int FINE_GRID = kSMs * 8;

// Fine-level operator of a hierarchy inside the cycle (cycle-precision matrix copy and vectors, fp64 right-hand side):
// out = b - A in (mode 1) or one damped Jacobi sweep (mode 2, optionally with the partials of b.out in mg.partial).
// dotSlot >= 0 (mode 2): b.out is reduced into mg.scal[dotSlot] by the same launch.
// MOF_MG_LASTFOLD=0 goes back to one k_fold launch per reduction (for A/B timing).
bool last_cta_folds() {
    static const bool on = env_int("MOF_MG_LASTFOLD", 1) != 0;
    return on;
}
Fold fold_into(Multigrid& mg, int slot) { return Fold{mg.partial.p, last_cta_folds() ? slot : -1, mg.scal.p, mg.counter.p}; }
int fold_after(mof_ctx* ctx, Multigrid& mg, int np, int slot) {
    if (!last_cta_folds()) MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, np, slot, mg.scal.p);
    return MOF_OK;
}
constexpr Fold NO_FOLD = {nullptr, -1, nullptr, nullptr};

bool scalar_row_kernel() {
    static const bool on = env_int("MOF_SCALAR_ROWKERNEL", 1) != 0;
    return on;
}
// MOF_SCALAR_SELL=0: the V x V operators stay in CSR inside the solver (A/B timing; the partitioned path always does)
bool scalar_sell(const mof_ctx* ctx) {
    static const bool on = env_int("MOF_SCALAR_SELL", 1) != 0;
    return on && ctx->sPadded > 0;
}

int fine_apply(mof_ctx* ctx, Multigrid& mg, const double* b, const double* omega, const creal* in, creal* out, int mode, int dotSlot = -1) {
    const Fold f = dotSlot >= 0 ? fold_into(mg, dotSlot) : NO_FOLD;
    if (mg.kind != MG_FLOW && scalar_sell(ctx)) {
        MOF_LAUNCH((k_fine_apply_scalar_sell<creal, creal>), FINE_GRID, B, 0, ctx->V, ctx->sSliceBase.p, ctx->sColSell.p, mg.fvalSell.p, b, mg.fdinv.p, omega, in, out, mode, f);
        return MOF_OK;
    }
    if (mg.kind != MG_FLOW && scalar_row_kernel()) {
        MOF_LAUNCH((k_fine_apply_scalar_row<creal, creal>), FINE_GRID, B, 0, ctx->V, ctx->sRowptr.p, ctx->sCol.p, mg.fval.p, b, mg.fdinv.p, omega, in, out, mode, f);
        return MOF_OK;
    }
    if (mg.kind == MG_FLOW)
        MOF_LAUNCH((k_fine_apply_flow<creal, creal>), FINE_GRID, B, 0, ctx->E, ctx->wSliceBase.p, ctx->wCol.p, mg.fval.p, b, mg.fdinv.p, omega, in, out, mode, f, 0, 0);
    else
        MOF_LAUNCH((k_fine_apply_scalar<creal, creal>), FINE_GRID, B, 0, ctx->V, ctx->sRowptr.p, ctx->sCol.p, mg.fval.p, b, mg.fdinv.p, omega, in, out, mode, f, 0, 0);
    return MOF_OK;
}
// out = b - A in with the fp64 matrix (initial and true residuals of PCG)
int fine_residual(mof_ctx* ctx, Multigrid& mg, const double* b, const double* in, double* out) {
    if (mg.kind == MG_FLOW)
        MOF_LAUNCH((k_fine_apply_flow<double, double>), FINE_GRID, B, 0, ctx->E, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, b, (const creal*)nullptr, mg.om(OM_ZERO), in, out, 1,
                   NO_FOLD, 0, 0);
    else if (scalar_sell(ctx))
        MOF_LAUNCH((k_fine_apply_scalar_sell<double, double>), FINE_GRID, B, 0, ctx->V, ctx->sSliceBase.p, ctx->sColSell.p, ctx->sSysSell.p, b, (const creal*)nullptr,
                   mg.om(OM_ZERO), in, out, 1, NO_FOLD);
    else
        MOF_LAUNCH((k_fine_apply_scalar<double, double>), FINE_GRID, B, 0, ctx->V, ctx->sRowptr.p, ctx->sCol.p, ctx->sSys.p, b, (const creal*)nullptr, mg.om(OM_ZERO), in, out, 1,
                   NO_FOLD, 0, 0);
    return MOF_OK;
}

// `zc` (with the level's parent table) = a coarse correction still to be added to lv.z, see k_coarse_apply.
template <int K, int D>
int coarse_apply(mof_ctx* ctx, MgLevel& lv, const double* omega, int mode, creal* out, const creal* zc) {
    const int* parent = zc ? lv.parent.p : nullptr;
    static const int wideFrom = env_int("MOF_MG_WIDE_FROM", 16384);  // cells from which the 9-warp variant (three slots per warp) replaces the 27-warp one
    if (lv.N >= wideFrom)
        MOF_LAUNCH((k_coarse_apply<K, D, 3>), blocks_for(lv.N, 32), 9 * 32, 0, lv.cblocks.p, lv.nbr.p, lv.binv.p, lv.r.p, lv.z.p, omega, lv.N, mode, out, parent, zc, 0, -1);
    else
        MOF_LAUNCH((k_coarse_apply<K, D, 1>), blocks_for(lv.N, 32), 27 * 32, 0, lv.cblocks.p, lv.nbr.p, lv.binv.p, lv.r.p, lv.z.p, omega, lv.N, mode, out, parent, zc, 0, -1);
    return MOF_OK;
}
int coarse_apply(mof_ctx* ctx, Multigrid& mg, MgLevel& lv, const double* omega, int mode, creal* out, const creal* zc = nullptr) {
    return mg.kind == MG_FLOW ? coarse_apply<9, 3>(ctx, lv, omega, mode, out, zc) : coarse_apply<1, 6>(ctx, lv, omega, mode, out, zc);
}

// Value-dependent part, once per system: coarser Galerkin levels, inverses, damping factors, dense coarsest inverse.
// The level-1 coefficients must already be in mg.lev[0].blocks.
int finish_values(mof_ctx* ctx, Multigrid& mg) {
    const int D = mg.dofs();
    const size_t len = mg.fineLen();
    for (int l = 0; l + 1 < mg.K; l++) {
        MgLevel& lv = mg.lev[l];
        MgLevel& up = mg.lev[l + 1];
        if (mg.kind == MG_FLOW)
            MOF_LAUNCH(k_coarsen_blocks<9>, blocks_for(27ll * up.N, B), B, 0, up.firstChild.p, lv.nbr.p, lv.parent.p, lv.blocks.p, lv.N, up.nbr.p, up.N, coarse_scale(mg, l + 1), up.blocks.p);
        else
            MOF_LAUNCH(k_coarsen_blocks<1>, blocks_for(27ll * up.N, B), B, 0, up.firstChild.p, lv.nbr.p, lv.parent.p, lv.blocks.p, lv.N, up.nbr.p, up.N, coarse_scale(mg, l + 1), up.blocks.p);
    }
    // Damping of the Jacobi smoothers: omega_l = 1.4 / rho_l with rho_l the spectral radius of Minv A on that level,
    // estimated by power iteration on I + Minv A (all eigenvalues of Minv A are positive). omega * rho < 2 keeps the
    // cycle positive definite; should an estimate ever be too low, PCG stalls and the caller falls back to Jacobi-PCG.
    // The iterate of the previous system on this mesh is kept (Multigrid::eig): consecutive systems differ little (the data term moves
    // with the flow, eps shrinks by 4), so from the second system on ONE step from there replaces 10 from a pseudo-random start (3 until the end of round 2: 542 -> 530 ms per
    // alignment for the same iteration counts; the estimate approaches rho from below and omega = 1.4 / rho leaves a factor 1.43 before omega rho = 2) —
    // the estimates only improve, and the set-up of the ~20 systems of an alignment drops from ~45 ms to ~20 ms at 1M vertices.
    const bool warm = mg.eigValid && env_int("MOF_MG_POWER_WARM", 1) != 0;
    const int powerIts = std::max(1, std::min(50, warm ? env_int("MOF_MG_POWER_ITS_WARM", 1) : env_int("MOF_MG_POWER_ITS", 10)));
    const char* pinvEnv = getenv("MOF_MG_PINV_TOL");
    const double pinvTol = pinvEnv && *pinvEnv ? atof(pinvEnv) : 1e-3;
    size_t eigOffset = 0;
    {
        MOF_CUDA(mg.eig.alloc(len + (size_t)D * [&] { size_t n = 0; for (int l = 0; l < mg.K; l++) n += mg.lev[l].N; return n; }()));
        MOF_CUDA(cudaMemsetAsync(mg.fq.p, 0, sizeof(double) * len, ctx->stream));  // the zero right-hand side
        creal* v = mg.fz.p;
        creal* w = mg.fz2.p;
        if (warm) MOF_CUDA(cudaMemcpyAsync(v, mg.eig.p, sizeof(creal) * len, cudaMemcpyDeviceToDevice, ctx->stream));
        else MOF_LAUNCH(k_pseudo_random, blocks_for((long long)len, B), B, 0, (long long)len, v);
        for (int it = 0; it <= powerIts; it++) {
            MOF_LAUNCH((k_dot_partial<creal, creal>), NBLK, B, 0, v, v, (long long)len, mg.partial.p);
            MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, 16, mg.scal.p);
            if (it == powerIts) break;
            MOF_LAUNCH(k_normalise, blocks_for((long long)len, B), B, 0, v, mg.scal.p, 16, (long long)len, v);
            MOF_TRY(fine_apply(ctx, mg, mg.fq.p, mg.om(OM_MINUS_ONE), v, w, 2));
            std::swap(v, w);
        }
        MOF_CUDA(cudaMemcpyAsync(mg.eig.p, v, sizeof(creal) * len, cudaMemcpyDeviceToDevice, ctx->stream));
        eigOffset = len;
    }
    for (int l = 0; l < mg.K; l++) {
        MgLevel& lv = mg.lev[l];
        MOF_LAUNCH(k_to_creal, kSMs * 4, B, 0, lv.blocks.p, (long long)lv.blocks.n, lv.cblocks.p);
        if (mg.kind == MG_FLOW) MOF_LAUNCH(k_block_pinv, blocks_for(lv.N, B), B, 0, lv.blocks.p, lv.N, pinvTol, lv.binv.p);
        else MOF_LAUNCH(k_scalar_inv, blocks_for(lv.N, B), B, 0, lv.blocks.p, lv.N, lv.binv.p);
        if (l == mg.K - 1) break;
        const long long nd = (long long)D * lv.N;
        MOF_CUDA(cudaMemsetAsync(lv.r.p, 0, sizeof(creal) * nd, ctx->stream));
        creal* const z0 = lv.z.p;  // coarse_apply reads lv.z: the iterate has to be there, so the two buffers trade places and are put back
        creal* const t0 = lv.t.p;
        if (warm) MOF_CUDA(cudaMemcpyAsync(lv.z.p, mg.eig.p + eigOffset, sizeof(creal) * nd, cudaMemcpyDeviceToDevice, ctx->stream));
        else MOF_LAUNCH(k_pseudo_random, blocks_for(nd, B), B, 0, nd, lv.z.p);
        for (int it = 0; it <= powerIts; it++) {
            MOF_LAUNCH((k_dot_partial<creal, creal>), NBLK, B, 0, lv.z.p, lv.z.p, nd, mg.partial.p);
            MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, 17 + l, mg.scal.p);
            if (it == powerIts) break;
            MOF_LAUNCH(k_normalise, blocks_for(nd, B), B, 0, lv.z.p, mg.scal.p, 17 + l, nd, lv.z.p);
            MOF_TRY(coarse_apply(ctx, mg, lv, mg.om(OM_MINUS_ONE), 2, lv.t.p));
            std::swap(lv.z.p, lv.t.p);
        }
        MOF_CUDA(cudaMemcpyAsync(mg.eig.p + eigOffset, lv.z.p, sizeof(creal) * nd, cudaMemcpyDeviceToDevice, ctx->stream));
        lv.z.p = z0, lv.t.p = t0;
        eigOffset += nd;
    }
    mg.eigValid = true;
    // coarsest level: dense matrix on the host, Cholesky inverse, back to the device
    MgLevel& lc = mg.lev.back();
    const int Kc = mg.comps(), bs = mg.kind == MG_FLOW ? 3 : 1;
    const int Nc = lc.N, nc = bs * Nc;
    mg.hostBlocks.resize((size_t)Nc * 27 * Kc);
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
    for (int l = 0; l + 1 < mg.K; l++) mg.lev[l].omega = damping(hs[17 + l]);
    MOF_TRY(upload_omegas(ctx, mg));
    if (env_int("MOF_MG_VERBOSE", 0)) {
        fprintf(stderr, "[mg %s] fine: n=%d rho %.3f omega %.3f\n", mg.kind == MG_FLOW ? "flow" : "scalar", mg.nFine, std::sqrt(hs[16]) - 1., mg.omega0);
        for (int l = 0; l < mg.K; l++)
            fprintf(stderr, "[mg %s] level %d: grid 2^%d, %d cells, rho %.3f omega %.3f\n", mg.kind == MG_FLOW ? "flow" : "scalar", l + 1, mg.lev[l].gridLevel, mg.lev[l].N,
                    l + 1 < mg.K ? std::sqrt(hs[17 + l]) - 1. : 0., mg.lev[l].omega);
    }
    std::vector<double>& M = mg.hostDense;
    M.assign((size_t)nc * nc, 0.);
    for (int I = 0; I < Nc; I++)
        for (int s = 0; s < 27; s++) {
            int J = mg.hostNbr[(size_t)I * 27 + s];
            if (J < 0) continue;
            for (int r = 0; r < bs; r++)
                for (int c = 0; c < bs; c++) {
                    size_t idx = ((size_t)s * Kc + (bs * r + c)) * Nc + I;
                    M[(size_t)(bs * I + r) * nc + bs * J + c] += mg.hostBlocks[idx];
                }
        }
    double trace = 0;
    for (int i = 0; i < nc; i++) trace += M[(size_t)i * nc + i];
    for (int i = 0; i < nc; i++) {
        M[(size_t)i * nc + i] += 1e-12 * trace / nc;
        for (int j = 0; j < i; j++) M[(size_t)i * nc + j] = M[(size_t)j * nc + i] = 0.5 * (M[(size_t)i * nc + j] + M[(size_t)j * nc + i]);
    }
    // Cholesky M = L L^T, then inverse = L^-T L^-1
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
    std::vector<double> Li((size_t)nc * nc, 0.);
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
    std::vector<creal> Mc(M.begin(), M.end());
    MOF_CUDA(cudaMemcpyAsync(mg.cinv.p, Mc.data(), sizeof(creal) * (size_t)nc * nc, cudaMemcpyHostToDevice, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    return MOF_OK;
}

bool fuse_residual_restrict() { return env_int("MOF_MG_FUSE_RESTRICT", 1) != 0; }  // (0: the two stand-alone kernels, for A/B timing)

int cycle_passes(const Multigrid& mg, int l) {
    const int wLevels = mg.gammaLevels >= 0 ? mg.gammaLevels : mg.fullDepth - (mg.kind == MG_FLOW ? 4 : 5);
    return l < wLevels ? mg.gamma : 1;
}

// The operations of a visit of level l (see coarse_cycle), appended to a k_coarse_tail program. `flip` follows which of a
// level's two buffers holds its iterate.
struct TailProgram {
    TailArgs a;
    int flip[MAXL + 2];
    bool overflow = false;
    void emit(unsigned char op, int l) {
        if (a.nOps >= TAIL_MAX_OPS - 1) { overflow = true; return; }
        a.op[a.nOps] = op, a.lvl[a.nOps] = (unsigned char)l, a.flip[a.nOps] = (unsigned char)(flip[l] | (flip[l + 1] << 1));
        a.nOps++;
    }
    void restrict_from(const Multigrid& mg, int l) { emit(l + 1 == mg.K - 1 ? T_DENSE : T_RESTRICT, l); }
    void visit(const Multigrid& mg, int l) {
        if (l == mg.K - 1) return;
        const int passes = cycle_passes(mg, l);
        for (int g = 0; g < passes; g++) {
            emit(T_RESIDUAL, l);
            restrict_from(mg, l);
            visit(mg, l + 1);
            if (g + 1 < passes) emit(T_PROLONG, l);  // W: the correction has to be in the iterate before the next residual
        }
        emit(T_SMOOTH, l);
        flip[l] ^= 1;
    }
};

// Shared-memory layout of a launch that holds levels first .. K-1 on a cluster of `ctas` CTAs; fills the level table of `a`
// and returns the bytes of dynamic shared memory per CTA.
size_t tail_layout(const Multigrid& mg, int first, int ctas, int residentCells, TailArgs& a) {
    const int K = mg.comps(), D = mg.dofs(), bs = mg.kind == MG_FLOW ? 3 : 1;
    int words = 0;
    auto take = [&](int n) { int o = words; words += (n + 3) & ~3; return o; };
    a.first = first, a.last = mg.K - 1;
    a.nDense = bs * mg.lev.back().N;
    for (int l = 0; l < mg.K; l++) {
        const MgLevel& lv = mg.lev[l];
        TailLevel& t = a.lev[l];
        t.blocks = lv.cblocks.p, t.nbr = lv.nbr.p, t.binv = lv.binv.p, t.parent = lv.parent.p, t.firstChild = lv.firstChild.p;
        t.r = lv.r.p, t.z = lv.z.p, t.t = lv.t.p, t.N = lv.N;
        t.cpc = (lv.N + ctas - 1) / ctas;
        t.resident = lv.N <= residentCells;
        if (l < first) continue;
        const int n = t.cpc;
        if (l == mg.K - 1) {
            t.oChild = take(lv.N + 1);
            t.oBuf[0] = t.oBuf[1] = take(D * n);
            a.oCinv = take(bs * n * a.nDense);
            continue;
        }
        t.oNbr = take(27 * n), t.oBinv = take(K * n), t.oParent = take(n), t.oChild = take(n + 1);
        t.oR = take(D * n), t.oBuf[0] = take(D * n), t.oBuf[1] = take(D * n);
        t.oBlocks = t.resident ? take(27 * K * n) : 0;
    }
    a.oPart = take(27 * TAIL_BATCH * D), a.oRes = take(TAIL_BATCH * D), a.oDense = take(DENSE_MAX);
    a.smemBytes = words * (int)sizeof(creal);
    return (size_t)a.smemBytes;
}

int tail_resident_cells() { return env_int("MOF_MG_TAIL_RESIDENT", 1536); }

// Everything from level `first` down in one launch: with fromAbove, the restriction of lev[first - 1].t comes first; without,
// lev[first].r and the pre-smoothed lev[first].z are read from global memory. The result is written to lev[first].z.
int launch_tail(mof_ctx* ctx, Multigrid& mg, int first, bool fromAbove) {
    TailProgram prog;
    TailArgs& a = prog.a;
    memset(&a, 0, sizeof(a));
    for (int l = 0; l <= MAXL + 1; l++) prog.flip[l] = 0;
    const size_t smem = tail_layout(mg, first, mg.tailGrid, tail_resident_cells(), a);
    a.cinv = mg.cinv.p, a.omega = mg.domega.p;
    a.trace = mg.tailTrace.p;  // nullptr unless MOF_MG_TAIL_TRACE
    if (fromAbove) prog.restrict_from(mg, first - 1);
    else prog.emit(T_LOAD, first);
    prog.visit(mg, first);
    if (prog.overflow) return fail(ctx, MOF_E_INVALID, "multigrid: the program of the small levels does not fit");
    a.flip[a.nOps] = (unsigned char)prog.flip[first];
#ifdef MOF_HOST_EMULATION
    (void)smem;
    if (mg.kind == MG_FLOW) mof_emul::submit_cooperative(mg.tailGrid, TAIL_T, [a] { k_coarse_tail<9, 3>(a); });
    else mof_emul::submit_cooperative(mg.tailGrid, TAIL_T, [a] { k_coarse_tail<1, 6>(a); });
    cudaError_t e = cudaSuccess;
#else
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(mg.tailGrid), cfg.blockDim = dim3(TAIL_T), cfg.dynamicSmemBytes = smem, cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = mg.tailGrid, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cudaError_t e = mg.kind == MG_FLOW ? cudaLaunchKernelEx(&cfg, k_coarse_tail<9, 3>, a) : cudaLaunchKernelEx(&cfg, k_coarse_tail<1, 6>, a);
#endif
    ctx->stats.kernelLaunches++;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaLaunchKernelEx(k_coarse_tail)");
    mg.traceN = a.nOps;
    for (int i = 0; i < a.nOps; i++) mg.traceOps[i] = (unsigned char)(a.op[i] * 16 + a.lvl[i]);
    return MOF_OK;
}

// The cluster: the largest the device co-schedules with this much shared memory — 16 CTAs (sm_100a, opt-in size), else 8, ...
template <int K, int D>
int tail_cluster_size(mof_ctx* ctx, size_t* smemLimit, int* ctas) {
    *ctas = 0;
#ifdef MOF_HOST_EMULATION
    *smemLimit = (size_t)64 << 20;  // (one emulated CTA holds what sixteen would share)
    MOF_CUDA(cudaDeviceGetAttribute(ctas, cudaDevAttrMultiProcessorCount, ctx->device));  // one emulated CTA; three (OS threads) in the threads build
#else
    int optin = 0;
    MOF_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    cudaFuncAttributes fa;
    MOF_CUDA(cudaFuncGetAttributes(&fa, k_coarse_tail<K, D>));
    const size_t smem = (size_t)optin - fa.sharedSizeBytes - 256;  // what is left beside the kernel's static shared memory
    *smemLimit = smem;
    MOF_CUDA(cudaFuncSetAttribute(k_coarse_tail<K, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool wide = cudaFuncSetAttribute(k_coarse_tail<K, D>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    cudaGetLastError();
    const int want = std::max(2, std::min(TAIL_MAX_CTAS, env_int("MOF_MG_TAIL_CTAS", 16)));
    for (int c = wide ? want : std::min(want, 8); c >= 2 && !*ctas; c /= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c), cfg.blockDim = dim3(TAIL_T), cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = c, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, k_coarse_tail<K, D>, &cfg) == cudaSuccess && clusters >= 1) *ctas = c;
        cudaGetLastError();
    }
#endif
    return MOF_OK;
}

// OFF by default (MOF_MG_TAIL_CELLS=0): on a B200 the launch is no faster than the 7-13 stand-alone kernels it replaces
// (1M vertices: 626 us per flow PCG iteration without it, 651 us with levels 3-5 on the cluster, 740 us with level 2 streamed as
// well) — staging the tables costs 10-18 us per launch and an operation still takes 1.4-3 us (cluster barrier with release /
// acquire ~0.7 us, distributed-shared-memory gathers at ~20 B/cycle/SM); profiles/r2f_tail_dsmem_trace.txt. Kept as an
// experiment that is bit-identical to the stand-alone path and covered by the CPU tier (MOF_MG_TAIL_CELLS=6144 there).
// Which levels go into k_coarse_tail: from the first one with at most MOF_MG_TAIL_CELLS cells (0: none) whose
// layout — coefficients resident up to MOF_MG_TAIL_RESIDENT cells (default 1536), streamed above — fits the shared memory of
// the cluster, down to the dense coarsest level. The dense level alone is not worth it (it already is one kernel).
int plan_tail(mof_ctx* ctx, Multigrid& mg) {
    mg.tailStart = -1;
    const int cells = env_int("MOF_MG_TAIL_CELLS", 0);
    if (cells <= 0 || mg.K < 2) return MOF_OK;
    size_t limit = 0;
    if (mg.kind == MG_FLOW) MOF_TRY((tail_cluster_size<9, 3>(ctx, &limit, &mg.tailGrid)));
    else MOF_TRY((tail_cluster_size<1, 6>(ctx, &limit, &mg.tailGrid)));
    if (mg.tailGrid < 1) return MOF_OK;
    size_t bytes = 0;
    for (int l = 0; l + 1 < mg.K && mg.tailStart < 0; l++) {
        if (mg.lev[l].N > cells || mg.lev[l].N > 60000 * mg.tailGrid) continue;  // (a packed cell index has 16 bits for the index in the owner's range)
        TailProgram prog;
        memset(&prog.a, 0, sizeof(prog.a));
        for (int q = 0; q <= MAXL + 1; q++) prog.flip[q] = 0;
        bytes = tail_layout(mg, l, mg.tailGrid, tail_resident_cells(), prog.a);
        if (bytes > limit) continue;
        prog.emit(T_RESTRICT, 0);
        prog.visit(mg, l);
        if (!prog.overflow) mg.tailStart = l;
    }
    if (env_int("MOF_MG_VERBOSE", 0))
        fprintf(stderr, "[mg %s] small levels: from level %d down on one cluster of %d CTAs, %zu bytes of shared memory each\n", mg.kind == MG_FLOW ? "flow" : "scalar",
                mg.tailStart + 1, mg.tailGrid, bytes);
    return MOF_OK;
}

// One cycle on the coarse hierarchy below level l: lev[l].r and the pre-smoothed lev[l].z in (the restriction kernels
// leave both), lev[l].z out. Launches per level: residual, restriction (+ first sweep of the next level, or the dense
// solve on the coarsest), post-smoothing (which adds the coarse correction while gathering). From level mg.tailStart down
// it is one launch per visit (k_coarse_tail), which starts with the restriction from the level above.
int coarse_cycle(mof_ctx* ctx, Multigrid& mg, int l) {
    MgLevel& lv = mg.lev[l];
    const bool flow = mg.kind == MG_FLOW;
    if (l == mg.K - 1) return MOF_OK;  // solved by the restriction that filled it
    if (mg.tailStart >= 0 && l >= mg.tailStart) return launch_tail(ctx, mg, l, false);
    MgLevel& up = mg.lev[l + 1];
    const bool upDense = l + 1 == mg.K - 1;
    const int passes = cycle_passes(mg, l);
    const bool fused = !upDense && mg.tailStart != l + 1 && lv.N < env_int("MOF_MG_FUSE_BELOW", 16384) && lv.gridLevel - up.gridLevel == 1 && fuse_residual_restrict();
    const int sweeps = l < (int)(sizeof(mg.coarseSweeps) / sizeof(int)) ? mg.coarseSweeps[l] : 1;
    for (int extra = 1; extra < sweeps; extra++) {  // further pre-smoothing sweeps (the first is the restriction's)
        MOF_TRY(coarse_apply(ctx, mg, lv, mg.om(1 + l), 2, lv.t.p));
        std::swap(lv.z.p, lv.t.p);
    }
    for (int g = 0; g < passes; g++) {
        if (fused) {  // residual and restriction in one launch (small levels)
            if (flow)
                MOF_LAUNCH((k_residual_restrict<9, 3>), blocks_for(up.N, FUSE_G), 27 * 32, 0, lv.cblocks.p, lv.nbr.p, lv.r.p, lv.z.p, lv.N, up.firstChild.p, up.binv.p, mg.om(2 + l),
                           up.N, up.r.p, up.z.p);
            else
                MOF_LAUNCH((k_residual_restrict<1, 6>), blocks_for(up.N, FUSE_G), 27 * 32, 0, lv.cblocks.p, lv.nbr.p, lv.r.p, lv.z.p, lv.N, up.firstChild.p, up.binv.p, mg.om(2 + l),
                           up.N, up.r.p, up.z.p);
            MOF_TRY(coarse_cycle(ctx, mg, l + 1));
        } else {
        MOF_TRY(coarse_apply(ctx, mg, lv, mg.om(1 + l), 1, lv.t.p));
        if (mg.tailStart == l + 1) MOF_TRY(launch_tail(ctx, mg, l + 1, true));
        else {
            if (upDense) {
                const int n = (flow ? 3 : 1) * up.N;
                if (flow) MOF_LAUNCH(k_dense_restrict_apply<3>, DENSE_CTAS, B, 0, up.firstChild.p, lv.t.p, mg.cinv.p, n, up.z.p);
                else MOF_LAUNCH(k_dense_restrict_apply<6>, DENSE_CTAS, B, 0, up.firstChild.p, lv.t.p, mg.cinv.p, n, up.z.p);
            } else if (flow)
                MOF_LAUNCH((k_restrict_coarse<9, 3>), blocks_for(up.N, 128), 128, 0, up.firstChild.p, lv.t.p, up.N, up.binv.p, mg.om(2 + l), up.r.p, up.z.p, 0, -1);
            else
                MOF_LAUNCH((k_restrict_coarse<1, 6>), blocks_for(up.N, 128), 128, 0, up.firstChild.p, lv.t.p, up.N, up.binv.p, mg.om(2 + l), up.r.p, up.z.p, 0, -1);
            MOF_TRY(coarse_cycle(ctx, mg, l + 1));
        }
        }
        if (g + 1 < passes) {  // W-cycle: the correction has to be in z before the next residual
            if (flow) MOF_LAUNCH(k_prolong_coarse<3>, blocks_for(3ll * lv.N, B), B, 0, lv.parent.p, up.z.p, lv.N, lv.z.p, 0, -1);
            else MOF_LAUNCH(k_prolong_coarse<6>, blocks_for(6ll * lv.N, B), B, 0, lv.parent.p, up.z.p, lv.N, lv.z.p, 0, -1);
        }
    }
    MOF_TRY(coarse_apply(ctx, mg, lv, mg.om(1 + l), 2, lv.t.p, up.z.p));
    std::swap(lv.z.p, lv.t.p);
    for (int extra = 1; extra < sweeps; extra++) {  // ... and as many more after the correction (symmetry)
        MOF_TRY(coarse_apply(ctx, mg, lv, mg.om(1 + l), 2, lv.t.p));
        std::swap(lv.z.p, lv.t.p);
    }
    return MOF_OK;
}

// z = cycle(r) on the fine level; result in mg.fz, and r.z folded into mg.scal[rzSlot]. With `presmoothed` the first
// sweep from a zero guess (z = omega0 * dinv * r) is already in mg.fz (k_update_xr leaves it there).
int fine_cycle(mof_ctx* ctx, Multigrid& mg, const double* r, bool presmoothed, int rzSlot) {
    const long long len = (long long)mg.fineLen();
    MgLevel& l1 = mg.lev[0];
    if (!presmoothed) MOF_LAUNCH(k_fine_presmooth<creal>, blocks_for(len, B), B, 0, r, mg.fdinv.p, mg.om(0), len, mg.nrhs, mg.fz.p);
    for (int sweep = 1; sweep < mg.fineSweeps; sweep++) {  // further pre-smoothing sweeps (the first one is the scaling of r above)
        MOF_TRY(fine_apply(ctx, mg, r, mg.om(0), mg.fz.p, mg.fz2.p, 2));
        std::swap(mg.fz.p, mg.fz2.p);
    }
    MOF_TRY(fine_apply(ctx, mg, r, mg.om(0), mg.fz.p, mg.ft.p, 1));
    if (mg.kind == MG_FLOW)
        MOF_LAUNCH(k_restrict_flow, blocks_for(32ll * l1.N, B), B, 0, mg.aggPtr.p, mg.aggList.p, mg.cevecAgg.p, mg.ft.p, l1.N, l1.binv.p, mg.om(1), l1.r.p, l1.z.p, 0, mg.nFine, 0, -1);
    else
        MOF_LAUNCH(k_restrict_scalar, blocks_for(6ll * l1.N, B), B, 0, mg.aggPtr.p, mg.aggList.p, mg.ft.p, l1.N, l1.binv.p, mg.om(1), l1.r.p, l1.z.p, 0, mg.nFine, 0, -1);
    if (mg.K == 1) {  // the aggregates are already the coarsest level
        const bool flow = mg.kind == MG_FLOW;
        const int n = (flow ? 3 : 1) * l1.N;
        if (flow) MOF_LAUNCH(k_dense_restrict_apply<3>, DENSE_CTAS, B, 0, (const int*)nullptr, l1.r.p, mg.cinv.p, n, l1.z.p);
        else MOF_LAUNCH(k_dense_restrict_apply<6>, DENSE_CTAS, B, 0, (const int*)nullptr, l1.r.p, mg.cinv.p, n, l1.z.p);
    }
    MOF_TRY(coarse_cycle(ctx, mg, 0));
    if (mg.kind == MG_FLOW) MOF_LAUNCH(k_prolong_flow, blocks_for(mg.nFine, B), B, 0, mg.agg.p, mg.cevec.p, l1.z.p, mg.nFine, mg.fz.p);
    else MOF_LAUNCH(k_prolong_scalar, blocks_for(len, B), B, 0, mg.agg.p, l1.z.p, mg.nFine, mg.fz.p);
    for (int sweep = 1; sweep < mg.fineSweeps; sweep++) {  // post-smoothing: as many sweeps as before the coarse correction (symmetry)
        MOF_TRY(fine_apply(ctx, mg, r, mg.om(0), mg.fz.p, mg.fz2.p, 2));
        std::swap(mg.fz.p, mg.fz2.p);
    }
    MOF_TRY(fine_apply(ctx, mg, r, mg.om(0), mg.fz.p, mg.fz2.p, 2, rzSlot));
    MOF_TRY(fold_after(ctx, mg, FINE_GRID, rzSlot));
    std::swap(mg.fz.p, mg.fz2.p);
    return MOF_OK;
}

// q = A p (fp64) with p.q folded into mg.scal[S_PQ] (and alpha derived). FLOW runs the PCG's own SpMV kernel of
// pcg_kernels.cu (the roofline kernel, left as it is) and folds its partials with one more launch.
int apply_dot(mof_ctx* ctx, Multigrid& mg, const double* p, double* q) {
    if (mg.kind == MG_FLOW) {
        int np = 0;
        MOF_TRY(spmv_dot_launch(ctx, ctx->E, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, p, q, mg.partial.p, &np));
        MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, np, S_PQ, mg.scal.p);
        return MOF_OK;
    }
    if (scalar_sell(ctx))
        MOF_LAUNCH((k_fine_apply_scalar_sell<double, double>), FINE_GRID, B, 0, ctx->V, ctx->sSliceBase.p, ctx->sColSell.p, ctx->sSysSell.p, (const double*)nullptr,
                   (const creal*)nullptr, mg.om(OM_ZERO), p, q, 0, fold_into(mg, S_PQ));
    else if (scalar_row_kernel())
        MOF_LAUNCH((k_fine_apply_scalar_row<double, double>), FINE_GRID, B, 0, ctx->V, ctx->sRowptr.p, ctx->sCol.p, ctx->sSys.p, (const double*)nullptr,
                   (const creal*)nullptr, mg.om(OM_ZERO), p, q, 0, fold_into(mg, S_PQ));
    else
        MOF_LAUNCH((k_fine_apply_scalar<double, double>), FINE_GRID, B, 0, ctx->V, ctx->sRowptr.p, ctx->sCol.p, ctx->sSys.p, (const double*)nullptr,
                   (const creal*)nullptr, mg.om(OM_ZERO), p, q, 0, fold_into(mg, S_PQ), 0, 0);
    return fold_after(ctx, mg, FINE_GRID, S_PQ);
}

// PCG with the cycle as preconditioner on the hierarchy's fine system: A x = b. With zeroGuess x starts at 0, otherwise
// from its content. SCALAR treats the six channels as one block-diagonal system (one alpha/beta for all), which keeps
// every vector operation flat; the stopping test is on the stacked residual.
int mg_pcg_replay(mof_ctx* ctx, Multigrid& mg, const double* b, double* x, bool zeroGuess, double tol, int maxIters, int* itersOut, double* relresOut) {
    const long long len = (long long)mg.fineLen();
    double* r = mg.fr.p;
    double* p = mg.fp.p;
    double* q = mg.fq.p;
    if (zeroGuess) {
        MOF_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * len, ctx->stream));
        MOF_CUDA(cudaMemcpyAsync(r, b, sizeof(double) * len, cudaMemcpyDeviceToDevice, ctx->stream));
    } else
        MOF_TRY(fine_residual(ctx, mg, b, x, r));
    MOF_LAUNCH((k_dot_partial<double, double>), NBLK, B, 0, b, b, len, mg.partial.p);
    MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_BB, mg.scal.p);
    MOF_LAUNCH((k_dot_partial<double, double>), NBLK, B, 0, r, r, len, mg.partial.p);
    MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_RR, mg.scal.p);
    double h2[2] = {0, 0};
    MOF_CUDA(cudaMemcpyAsync(h2, mg.scal.p + S_RR, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    MOF_CUDA(cudaStreamSynchronize(ctx->stream));
    double rr = h2[0];
    const double bb = h2[1];
    *itersOut = 0, *relresOut = 0;
    if (!(bb > 0)) {
        if (!zeroGuess) MOF_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * len, ctx->stream));
        return MOF_OK;
    }
    int it = 0;
    for (int attempt = 0; attempt < 4 && rr > tol * tol * bb; attempt++) {
        // (re)start: z = M r, p = z, rz = r.z
        MOF_TRY(fine_cycle(ctx, mg, r, false, S_RZ));
        MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p, mg.scal.p, len, 1, p);
        // One PCG iteration is ~35 small dependent launches (most of them on the tiny coarse levels): capture TWO
        // iterations once as a CUDA graph and replay it (the ping-pong buffers of the cycle are back in place after an
        // even number of cycles). The residual norms of both iterations land in pinned host memory.
        auto iteration = [&](int slot) -> int {
            MOF_TRY(apply_dot(ctx, mg, p, q));
            MOF_LAUNCH(k_update_xr, NBLK, B, 0, p, q, len, x, r, mg.fdinv.p, mg.om(0), mg.nrhs, mg.fz.p, fold_into(mg, S_RR));
            MOF_TRY(fold_after(ctx, mg, NBLK, S_RR));
            MOF_CUDA(cudaMemcpyAsync(mg.hostRR + slot, mg.scal.p + S_RR, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MOF_TRY(fine_cycle(ctx, mg, r, true, S_RZNEW));
            MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p, mg.scal.p, len, 0, p);
            return MOF_OK;
        };
        // (the path of MOF_MG_WHILE=0: a graph of two iterations per attempt, replayed from the host; mg_pcg below keeps the host out)
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
        MOF_TRY(fine_residual(ctx, mg, b, x, r));
        MOF_LAUNCH((k_dot_partial<double, double>), NBLK, B, 0, r, r, len, mg.partial.p);
        MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_RR, mg.scal.p);
        MOF_CUDA(cudaMemcpyAsync(&rr, mg.scal.p + S_RR, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MOF_CUDA(cudaStreamSynchronize(ctx->stream));
        if (it >= maxIters || !std::isfinite(rr)) break;
    }
    *itersOut = it;
    *relresOut = std::sqrt(rr / bb);
    if (!(*relresOut <= tol * 1.0001)) ctx->stats.solvesAboveTolerance++;
    if (!std::isfinite(rr) || (!(*relresOut <= tol * 1.0001) && (it >= maxIters || !(*relresOut <= MOF_ACCEPT_RELRES)))) {
        char msg[160];
        snprintf(msg, sizeof(msg), "[ERROR] multigrid PCG did not reach %g in %d iterations (relative residual %g)", tol, it, *relresOut);
        return fail(ctx, MOF_E_NOCONVERGE, msg);
    }
    return MOF_OK;
}

// The solve's graph for this pair of buffers: captured on first use, kept until the mesh changes.
//   [first cycle: z = M r, p = z, r.z] -> [continue?] -> WHILE { two PCG iterations ; continue? } -> [r = b - A x in fp64, r.r] -> [scalars to the host]
// The loop's condition is set on the device (k_pcg_continue), so a solve is one graph launch and one synchronisation, and the
// ping-pong buffers of the cycle are back in place after the two cycles of a pass.
int pcg_graph(mof_ctx* ctx, Multigrid& mg, const double* b, double* x, PcgGraph** out) {
    for (PcgGraph& g : mg.graphs)
        if (g.b == b && g.x == x) { *out = &g; return MOF_OK; }
    if (mg.graphs.size() >= 6) drop_graphs(mg);
    const long long len = (long long)mg.fineLen();
    double* r = mg.fr.p;
    double* p = mg.fp.p;
    double* q = mg.fq.p;
    PcgGraph g;
    g.b = b, g.x = x;
    const long long launchesBefore = ctx->stats.kernelLaunches;
    cudaGraph_t body = nullptr;
    cudaGraphConditionalHandle loop{};
    int crc = MOF_OK;
    cudaError_t ce = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed);
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "cudaStreamBeginCapture(pcg)");
    auto captured = [&]() -> int {
        MOF_TRY(fine_cycle(ctx, mg, r, false, S_RZ));
        MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p, mg.scal.p, len, 1, p);
        // the WHILE node goes where the capture stands
        cudaStreamCaptureStatus status;
        cudaGraph_t capturing = nullptr;
        const cudaGraphNode_t* deps = nullptr;
        size_t ndeps = 0;
        MOF_CUDA(cudaStreamGetCaptureInfo(ctx->stream, &status, nullptr, &capturing, &deps, &ndeps));
        MOF_CUDA(cudaGraphConditionalHandleCreate(&loop, capturing, 0, cudaGraphCondAssignDefault));
        MOF_LAUNCH(k_pcg_continue, 1, 1, 0, loop, mg.scal.p, 1);
        MOF_CUDA(cudaStreamGetCaptureInfo(ctx->stream, &status, nullptr, &capturing, &deps, &ndeps));
        cudaGraphNodeParams params = {cudaGraphNodeTypeConditional};
        params.type = cudaGraphNodeTypeConditional;
        params.conditional.handle = loop;
        params.conditional.type = cudaGraphCondTypeWhile;
        params.conditional.size = 1;
        cudaGraphNode_t node = nullptr;
        MOF_CUDA(cudaGraphAddNode(&node, capturing, deps, ndeps, &params));
        body = params.conditional.phGraph_out[0];
        MOF_CUDA(cudaStreamUpdateCaptureDependencies(ctx->stream, &node, 1, cudaStreamSetCaptureDependencies));
        // after the loop: the true residual of x and what the host wants to know
        MOF_TRY(fine_residual(ctx, mg, b, x, r));
        MOF_LAUNCH((k_dot_partial<double, double>), NBLK, B, 0, r, r, len, mg.partial.p);
        MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_RR, mg.scal.p);
        MOF_CUDA(cudaMemcpyAsync(mg.hostRR, mg.scal.p, sizeof(double) * S_REPORT, cudaMemcpyDeviceToHost, ctx->stream));
        return MOF_OK;
    };
    crc = captured();
    ce = cudaStreamEndCapture(ctx->stream, &g.graph);
    g.launchesOnce = ctx->stats.kernelLaunches - launchesBefore;
    if (crc == MOF_OK && ce == cudaSuccess && g.graph && body) {
        const long long before = ctx->stats.kernelLaunches;
        ce = cudaStreamBeginCaptureToGraph(ctx->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
        if (ce == cudaSuccess) {
            auto iteration = [&](int rrSlot) -> int {
                MOF_TRY(apply_dot(ctx, mg, p, q));
                MOF_LAUNCH(k_update_xr, NBLK, B, 0, p, q, len, x, r, mg.fdinv.p, mg.om(0), mg.nrhs, mg.fz.p, fold_into(mg, rrSlot));
                MOF_TRY(fold_after(ctx, mg, NBLK, rrSlot));
                MOF_TRY(fine_cycle(ctx, mg, r, true, S_RZNEW));
                MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p, mg.scal.p, len, 0, p);
                return MOF_OK;
            };
            auto pass = [&]() -> int {
                MOF_TRY(iteration(S_RR0));
                MOF_TRY(iteration(S_RR));
                MOF_LAUNCH(k_pcg_continue, 1, 1, 0, loop, mg.scal.p, 0);
                return MOF_OK;
            };
            crc = pass();
            cudaGraph_t same = nullptr;
            ce = cudaStreamEndCapture(ctx->stream, &same);
        }
        g.launchesPerPass = ctx->stats.kernelLaunches - before;
    }
    ctx->stats.kernelLaunches = launchesBefore;
    if (crc == MOF_OK && ce == cudaSuccess && g.graph) ce = cudaGraphInstantiate(&g.exec, g.graph, 0);
    if (crc != MOF_OK || ce != cudaSuccess || !g.exec) {
        if (g.graph) cudaGraphDestroy(g.graph);
        cudaGetLastError();
        return crc != MOF_OK ? crc : cuda_fail(ctx, ce, "capture of the PCG loop (conditional graph node)");
    }
    mg.graphs.push_back(g);
    *out = &mg.graphs.back();
    return MOF_OK;
}

bool pcg_while_enabled() { return env_int("MOF_MG_WHILE", 1) != 0; }  // (read per solve: tests flip it inside one process)

// PCG with the cycle as preconditioner on the hierarchy's fine system: A x = b. With zeroGuess x starts at 0, otherwise
// from its content. SCALAR treats the six channels as one block-diagonal system (one alpha/beta for all), which keeps
// every vector operation flat; the stopping test is on the stacked residual. One graph launch per attempt (pcg_graph); an
// attempt whose true residual misses the tolerance (drift of the recurrence) is followed by another from the x reached.
int mg_pcg(mof_ctx* ctx, Multigrid& mg, const double* b, double* x, bool zeroGuess, double tol, int maxIters, int* itersOut, double* relresOut) {
    if (!pcg_while_enabled() || mg.noWhile) return mg_pcg_replay(ctx, mg, b, x, zeroGuess, tol, maxIters, itersOut, relresOut);
    PcgGraph* g = nullptr;
    if (pcg_graph(ctx, mg, b, x, &g) != MOF_OK) {  // no conditional nodes on this driver: the replay loop does the same job
        mg.noWhile = true;
        return mg_pcg_replay(ctx, mg, b, x, zeroGuess, tol, maxIters, itersOut, relresOut);
    }
    const long long len = (long long)mg.fineLen();
    double* r = mg.fr.p;
    if (zeroGuess) {
        MOF_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * len, ctx->stream));
        MOF_CUDA(cudaMemcpyAsync(r, b, sizeof(double) * len, cudaMemcpyDeviceToDevice, ctx->stream));
    } else
        MOF_TRY(fine_residual(ctx, mg, b, x, r));
    MOF_LAUNCH((k_dot_partial<double, double>), NBLK, B, 0, b, b, len, mg.partial.p);
    MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_BB, mg.scal.p);
    MOF_LAUNCH((k_dot_partial<double, double>), NBLK, B, 0, r, r, len, mg.partial.p);
    MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, S_RR, mg.scal.p);
    MOF_LAUNCH(k_pcg_begin, 1, 1, 0, mg.scal.p, tol * tol, maxIters);
    *itersOut = 0, *relresOut = 0;
    double rr = 0, bb = 0;
    int it = 0;
    for (int attempt = 0; attempt < 4; attempt++) {
        cudaError_t ce = cudaGraphLaunch(g->exec, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
        if (ce != cudaSuccess) return cuda_fail(ctx, ce, "cudaGraphLaunch(pcg)");
        const double* h = mg.hostRR;
        const int itNow = (int)h[S_IT];
        ctx->stats.kernelLaunches += g->launchesOnce + g->launchesPerPass * ((itNow - it) / 2);
        it = itNow, rr = h[S_RR], bb = h[S_BB];
        if (!(bb > 0)) {
            if (!zeroGuess) MOF_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * len, ctx->stream));
            return MOF_OK;
        }
        if (!(rr > tol * tol * bb) || it >= maxIters || !std::isfinite(rr)) break;
    }
    if (mg.tailTrace.p && mg.traceN > 0) {  // diagnostic: where the time of the last k_coarse_tail launch went
        std::vector<unsigned long long> h(256);
        MOF_CUDA(read_back(ctx, h.data(), mg.tailTrace.p, h.size()));
        static const char* names[] = {"residual", "restrict", "dense", "prolong", "smooth", "load"};
        fprintf(stderr, "[mg tail %s] staging %.2f us, %d operations, %.2f us:", mg.kind == MG_FLOW ? "flow" : "scalar", (h[0] - h[255]) * 1e-3, mg.traceN,
                (h[mg.traceN] - h[0]) * 1e-3);
        for (int i = 0; i < mg.traceN; i++) fprintf(stderr, " %s%d=%.2f", names[mg.traceOps[i] / 16], mg.traceOps[i] % 16, (h[i + 1] - h[i]) * 1e-3);
        fprintf(stderr, "\n");
    }
    *itersOut = it;
    *relresOut = std::sqrt(rr / bb);
    if (!(*relresOut <= tol * 1.0001)) ctx->stats.solvesAboveTolerance++;
    if (!std::isfinite(rr) || (!(*relresOut <= tol * 1.0001) && (it >= maxIters || !(*relresOut <= MOF_ACCEPT_RELRES)))) {
        char msg[160];
        snprintf(msg, sizeof(msg), "[ERROR] multigrid PCG did not reach %g in %d iterations (relative residual %g)", tol, it, *relresOut);
        return fail(ctx, MOF_E_NOCONVERGE, msg);
    }
    return MOF_OK;
}

// ---------------------------------------------------------------------- one mesh over several GPUs (dist.cu)

inline int dist_halo(mof_ctx* c, int kind, float* v) { return dist_halo_f32(c, kind, v); }
inline int dist_halo(mof_ctx* c, int kind, double* v) { return dist_halo_f64(c, kind, v); }
inline int dist_allreduce(mof_ctx* c, float* v, int n) { return dist_allreduce_f32(c, v, n); }
inline int dist_allreduce(mof_ctx* c, double* v, int n) { return dist_allreduce_f64(c, v, n); }

// Raw slots of mg.scal for this rank's partial sums (all-reduced in place), and the scalar each one feeds.
enum { R_PQ = 8, R_RR = 9, R_RZNEW = 10, R_RZ = 11, R_BB = 12 };  // R_RR and R_RZNEW adjacent: one all-reduce carries both
__global__ void k_derive(int raw, double* __restrict__ scal) { pdl_wait();
    const double v = scal[raw];
    if (raw == R_PQ) scal[S_PQ] = v, scal[S_ALPHA] = v != 0 ? scal[S_RZ] / v : 0.;
    else if (raw == R_RR) scal[S_RR] = v;
    else if (raw == R_BB) scal[S_BB] = v;
    else if (raw == R_RZ) scal[S_RZ] = v;
    else if (raw == R_RZNEW) scal[S_BETA] = scal[S_RZ] != 0 ? v / scal[S_RZ] : 0., scal[S_RZ] = v;
}

// ---- coarse levels dealt to the ranks
// The replicated coarse levels are what limits a partitioned mesh (16.8M vertices: level 1 alone is 1.17M cells, 1.1 GB of stencil
// coefficients per sweep on EVERY rank, and its restricted residual a 14 MB all-reduce per cycle). Levels of more than
// MOF_DIST_LEVEL_CELLS cells (default 100 000) are therefore dealt to the ranks like the fine rows: contiguous cell ranges in
// Morton order, aligned through the octree (a rank's cells on level l are the children of its cells on level l + 1), so that
// restriction and prolongation between dealt levels stay inside a rank and only the 27-point stencils (one layer of cells) and
// the aggregates that straddle a row-block boundary need an exchange. The first replicated level is gathered once per visit.
__global__ void k_mark_stencil_halo(const int* __restrict__ nbr, int c0, int c1, int* __restrict__ flags) { pdl_wait();
    int i = 27 * c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 27 * c1) return;
    const int J = nbr[i];
    if (J >= 0 && (J < c0 || J >= c1)) flags[J] = 1;
}
__global__ void k_mark_aggregate_halo(const int* __restrict__ agg, int r0, int r1, int c0, int c1, int* __restrict__ flags) { pdl_wait();
    int e = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r1) return;
    const int a = agg[e];
    if (a < c0 || a >= c1) flags[a] = 1;
}
__global__ void k_mark_member_halo(const int* __restrict__ aggPtr, const int* __restrict__ aggList, int c0, int c1, int r0, int r1, int* __restrict__ flags) { pdl_wait();
    int q = aggPtr[c0] + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= aggPtr[c1]) return;
    const int e = aggList[q];
    if (e < r0 || e >= r1) flags[e] = 1;
}

int mg_dist_setup_one(mof_ctx* ctx, Multigrid& mg) {
    mg.distLevels = 0, mg.cellStart.clear(), mg.levelPart.clear(), mg.memberPart = mg.gatherPart = -1;
    const int world = dist_world(ctx), rank = dist_rank(ctx);
    if (!mg.usable || !std::is_same<creal, float>::value) return MOF_OK;
    // (the smoothing systems' levels stay replicated by default: on 8 B200s at 16.8M vertices dealing them cost 709 ms against 644 ms
    //  replicated — their cycle has one W level less and their exchanges are six values wide; profiles/r2o_partitioned_16M_8gpu.json)
    const int threshold = mg.kind == MG_FLOW ? env_int("MOF_DIST_LEVEL_CELLS", 100000) : env_int("MOF_DIST_LEVEL_CELLS_SCALAR", 0);
    int P = 0;
    while (P + 2 < mg.K && mg.lev[P].N > threshold) P++;  // (the level below the last dealt one is a stencil level: it is gathered, then replicated)
    if (P == 0 || threshold <= 0) return MOF_OK;
    const bool flow = mg.kind == MG_FLOW;
    const int kind = flow ? 0 : 1, D = mg.dofs();
    int s0, s1, r0, r1;
    dist_range(ctx, kind, &s0, &s1, &r0, &r1);
    // cell ranges, top down
    mg.cellStart.assign(P + 1, std::vector<int>(world + 1, 0));
    for (int k = 0; k <= world; k++) mg.cellStart[P][k] = (int)((long long)mg.lev[P].N * k / world);
    for (int l = P - 1; l >= 0; l--)
        for (int k = 0; k <= world; k++) MOF_CUDA(read_back(ctx, &mg.cellStart[l][k], mg.lev[l + 1].firstChild.p + mg.cellStart[l + 1][k]));
    size_t most = (size_t)mg.nFine + 1;
    for (int l = 0; l < P; l++) most = std::max(most, (size_t)mg.lev[l].N + 1);
    MOF_CUDA(ctx->itmp0.reserve(most));
    MOF_CUDA(ctx->itmp1.reserve(most));
    mg.levelPart.assign(P, -1);
    for (int l = 0; l < P; l++) {
        MgLevel& lv = mg.lev[l];
        const int c0 = mg.cellStart[l][rank], c1 = mg.cellStart[l][rank + 1];
        MOF_CUDA(cudaMemsetAsync(ctx->itmp0.p, 0, sizeof(int) * ((size_t)lv.N + 1), ctx->stream));
        if (c1 > c0) MOF_LAUNCH(k_mark_stencil_halo, blocks_for(27ll * (c1 - c0), B), B, 0, lv.nbr.p, c0, c1, ctx->itmp0.p);
        if (l == 0 && r1 > r0) MOF_LAUNCH(k_mark_aggregate_halo, blocks_for(r1 - r0, B), B, 0, mg.agg.p, r0, r1, c0, c1, ctx->itmp0.p);
        MOF_TRY(dist_add_partition(ctx, D, mg.cellStart[l].data(), lv.N, &mg.levelPart[l]));
    }
    {  // fine rows of my level-0 cells that other ranks own
        const int c0 = mg.cellStart[0][rank], c1 = mg.cellStart[0][rank + 1];
        int q0 = 0, q1 = 0;
        MOF_CUDA(read_back(ctx, &q0, mg.aggPtr.p + c0));
        MOF_CUDA(read_back(ctx, &q1, mg.aggPtr.p + c1));
        MOF_CUDA(cudaMemsetAsync(ctx->itmp0.p, 0, sizeof(int) * ((size_t)mg.nFine + 1), ctx->stream));
        if (q1 > q0) MOF_LAUNCH(k_mark_member_halo, blocks_for(q1 - q0, B), B, 0, mg.aggPtr.p, mg.aggList.p, c0, c1, r0, r1, ctx->itmp0.p);
        std::vector<int> rowStart(world + 1);
        dist_row_starts(ctx, kind, rowStart.data());  // the row blocks of this system, as dist.cu dealt them
        MOF_TRY(dist_add_partition(ctx, mg.nrhs, rowStart.data(), mg.nFine, &mg.memberPart));
    }
    {  // the first replicated level: every rank computes its range of it, all ranks get all of it
        MOF_CUDA(cudaMemsetAsync(ctx->itmp0.p, 0, sizeof(int) * ((size_t)mg.lev[P].N + 1), ctx->stream));
        MOF_TRY(dist_add_partition(ctx, D, mg.cellStart[P].data(), mg.lev[P].N, &mg.gatherPart));
    }
    mg.distLevels = P;
    drop_graphs(mg);
    if (env_int("MOF_MG_VERBOSE", 0)) {
        fprintf(stderr, "[mg %s] rank %d of %d: levels 1..%d dealt to the ranks;", flow ? "flow" : "scalar", rank, world, P);
        for (int l = 0; l < P; l++)
            fprintf(stderr, " level %d cells [%d, %d) of %d, halo %lld;", l + 1, mg.cellStart[l][rank], mg.cellStart[l][rank + 1], mg.lev[l].N, dist_partition_halo(ctx, mg.levelPart[l]));
        fprintf(stderr, " fine rows of my aggregates elsewhere: %lld\n", dist_partition_halo(ctx, mg.memberPart));
    }
    return MOF_OK;
}

// One visit of dealt level l (see coarse_cycle for the replicated form): lev[l].r and the pre-smoothed lev[l].z are valid on the own
// cells; so is lev[l].z afterwards.
int coarse_cycle_dist(mof_ctx* ctx, Multigrid& mg, int l) {
    if (l >= mg.distLevels) return coarse_cycle(ctx, mg, l);
    const bool flow = mg.kind == MG_FLOW;
    const int rank = dist_rank(ctx);
    MgLevel& lv = mg.lev[l];
    MgLevel& up = mg.lev[l + 1];
    const int c0 = mg.cellStart[l][rank], c1 = mg.cellStart[l][rank + 1], nc = c1 - c0;
    const int u0 = mg.cellStart[l + 1][rank], u1 = mg.cellStart[l + 1][rank + 1], nu = u1 - u0;
    auto apply = [&](int mode, creal* out) -> int {
        if (nc <= 0) return MOF_OK;
        if (flow) MOF_LAUNCH((k_coarse_apply<9, 3, 3>), blocks_for(nc, 32), 9 * 32, 0, lv.cblocks.p, lv.nbr.p, lv.binv.p, lv.r.p, lv.z.p, mg.om(1 + l), lv.N, mode, out, (const int*)nullptr,
                             (const creal*)nullptr, c0, c1);
        else MOF_LAUNCH((k_coarse_apply<1, 6, 3>), blocks_for(nc, 32), 9 * 32, 0, lv.cblocks.p, lv.nbr.p, lv.binv.p, lv.r.p, lv.z.p, mg.om(1 + l), lv.N, mode, out, (const int*)nullptr,
                        (const creal*)nullptr, c0, c1);
        return MOF_OK;
    };
    auto prolong = [&]() -> int {  // z += P zc on the own cells (their parents are own cells of the next level, or the next level is complete)
        if (nc <= 0) return MOF_OK;
        if (flow) MOF_LAUNCH(k_prolong_coarse<3>, blocks_for(3ll * nc, B), B, 0, lv.parent.p, up.z.p, lv.N, lv.z.p, c0, c1);
        else MOF_LAUNCH(k_prolong_coarse<6>, blocks_for(6ll * nc, B), B, 0, lv.parent.p, up.z.p, lv.N, lv.z.p, c0, c1);
        return MOF_OK;
    };
    const int passes = cycle_passes(mg, l);
    for (int g = 0; g < passes; g++) {
        MOF_TRY(dist_halo_part_f32(ctx, mg.levelPart[l], (float*)lv.z.p));
        MOF_TRY(apply(1, lv.t.p));
        if (nu > 0) {
            if (flow) MOF_LAUNCH((k_restrict_coarse<9, 3>), blocks_for(nu, 128), 128, 0, up.firstChild.p, lv.t.p, up.N, up.binv.p, mg.om(2 + l), up.r.p, up.z.p, u0, u1);
            else MOF_LAUNCH((k_restrict_coarse<1, 6>), blocks_for(nu, 128), 128, 0, up.firstChild.p, lv.t.p, up.N, up.binv.p, mg.om(2 + l), up.r.p, up.z.p, u0, u1);
        }
        if (l + 1 == mg.distLevels) {
            float* both[2] = {(float*)up.r.p, (float*)up.z.p};
            MOF_TRY(dist_allgather_part_f32(ctx, mg.gatherPart, both, 2));
        }
        MOF_TRY(coarse_cycle_dist(ctx, mg, l + 1));
        if (g + 1 < passes) MOF_TRY(prolong());
    }
    MOF_TRY(prolong());
    MOF_TRY(dist_halo_part_f32(ctx, mg.levelPart[l], (float*)lv.z.p));
    MOF_TRY(apply(2, lv.t.p));
    std::swap(lv.z.p, lv.t.p);
    return MOF_OK;
}

// The multigrid-PCG of either hierarchy with the fine level row-partitioned over the ranks. Every rank runs this with
// the same b (and the same initial x unless zeroGuess); vectors are full length, a rank computes the rows [r0, r1) of
// each (FLOW: = slices [s0, s1); SCALAR: six values per row). Per iteration: three halo exchanges (p in fp64, the
// cycle's iterate twice in fp32), three all-reduces (p.q; the level-1 restriction; r.r together with r.z — the residual
// norm only feeds the stopping test, so it waits for the cycle's own reduction); the coarse levels are replicated. x is
// complete on every rank at the end.
int mg_pcg_dist(mof_ctx* ctx, Multigrid& mg, const double* b, double* x, bool zeroGuess, double tol, int maxIters, int* itersOut, double* relresOut) {
    const bool flow = mg.kind == MG_FLOW;
    const int kind = flow ? 0 : 1, W = mg.nrhs;
    int s0, s1, r0, r1;
    dist_range(ctx, kind, &s0, &s1, &r0, &r1);
    const int rows = r1 - r0;
    const size_t e0 = (size_t)r0 * W;          // first element of my rows in a fine vector
    const long long len = (long long)rows * W;  // ... and their number
    const long long full = (long long)mg.fineLen();
    MgLevel& l1 = mg.lev[0];
    double* r = mg.fr.p;
    double* p = mg.fp.p;
    double* q = mg.fq.p;
    const Fold partials = {mg.partial.p, -1, mg.scal.p, mg.counter.p};
    // sum of this rank's partials -> all ranks' sum -> the PCG scalar it feeds
    auto reduce = [&](int np, int raw) -> int {
        MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, np, raw, mg.scal.p);
        MOF_TRY(dist_allreduce(ctx, mg.scal.p + raw, 1));
        MOF_LAUNCH(k_derive, 1, 1, 0, raw, mg.scal.p);
        return MOF_OK;
    };
    // my rows of the fine operator, on the cycle's fp32 copy (mg.fval) or on the fp64 system matrix
    const double* sysVal = flow ? ctx->wA.p : ctx->sSys.p;
    auto spmv = [&](const auto* val, const double* rhs, const double* omega, const auto* in, auto* out, int mode, Fold f) -> int {
        using TV = std::remove_cv_t<std::remove_pointer_t<decltype(val)>>;
        if (flow) MOF_LAUNCH((k_fine_apply_flow<TV, TV, true>), FINE_GRID, B, 0, ctx->E, ctx->wSliceBase.p, ctx->wCol.p, val, rhs, mg.fdinv.p, omega, in, out, mode, f, s0, s1);
        else MOF_LAUNCH((k_fine_apply_scalar<TV, TV, true>), FINE_GRID, B, 0, ctx->V, ctx->sRowptr.p, ctx->sCol.p, val, rhs, mg.fdinv.p, omega, in, out, mode, f, r0, r1);
        return MOF_OK;
    };
    const int D = mg.dofs();
    // z = cycle(r) on my rows (mg.fz), r.z into scal via `rawSlot`. With `withRR` this rank's part of r.r is waiting in
    // scal[R_RR] (folded by the caller): it joins the all-reduce of r.z (rawSlot = R_RZNEW, the slot next to it).
    auto cycle = [&](bool presmoothed, int rawSlot, bool withRR = false) -> int {
        if (!presmoothed && len) MOF_LAUNCH(k_fine_presmooth<creal>, blocks_for(len, B), B, 0, r + e0, mg.fdinv.p + r0, mg.om(0), len, W, mg.fz.p + e0);
        MOF_TRY(dist_halo(ctx, kind, mg.fz.p));
        MOF_TRY(spmv((const creal*)mg.fval.p, r, mg.om(0), (const creal*)mg.fz.p, mg.ft.p, 1, NO_FOLD));
        if (mg.distLevels > 0) {
            // level 1 is dealt to the ranks: a rank restricts ALL the rows of its own cells (those that other ranks own arrive by an
            // exchange of the residual), which is also the level's first sweep — no all-reduce over the level
            const int c0 = mg.cellStart[0][dist_rank(ctx)], c1 = mg.cellStart[0][dist_rank(ctx) + 1];
            MOF_TRY(dist_halo_part_f32(ctx, mg.memberPart, (float*)mg.ft.p));
            if (c1 > c0) {
                if (flow)
                    MOF_LAUNCH(k_restrict_flow, blocks_for(32ll * (c1 - c0), B), B, 0, mg.aggPtr.p, mg.aggList.p, mg.cevecAgg.p, mg.ft.p, l1.N, l1.binv.p, mg.om(1), l1.r.p, l1.z.p, 0,
                               mg.nFine, c0, c1);
                else
                    MOF_LAUNCH(k_restrict_scalar, blocks_for(6ll * (c1 - c0), B), B, 0, mg.aggPtr.p, mg.aggList.p, mg.ft.p, l1.N, l1.binv.p, mg.om(1), l1.r.p, l1.z.p, 0, mg.nFine,
                               c0, c1);
            }
            MOF_TRY(coarse_cycle_dist(ctx, mg, 0));
            MOF_TRY(dist_halo_part_f32(ctx, mg.levelPart[0], (float*)l1.z.p));  // the cells of my rows' aggregates that other ranks own
        } else {
        if (flow)
            MOF_LAUNCH(k_restrict_flow, blocks_for(32ll * l1.N, B), B, 0, mg.aggPtr.p, mg.aggList.p, mg.cevecAgg.p, mg.ft.p, l1.N, l1.binv.p, mg.om(1), l1.r.p, l1.z.p, r0, r1, 0, -1);
        else
            MOF_LAUNCH(k_restrict_scalar, blocks_for(6ll * l1.N, B), B, 0, mg.aggPtr.p, mg.aggList.p, mg.ft.p, l1.N, l1.binv.p, mg.om(1), l1.r.p, l1.z.p, r0, r1, 0, -1);
        MOF_TRY(dist_allreduce(ctx, l1.r.p, D * l1.N));
        if (flow) MOF_LAUNCH((k_level_presmooth<9, 3>), blocks_for(l1.N, B), B, 0, l1.binv.p, l1.r.p, mg.om(1), l1.N, l1.z.p);
        else MOF_LAUNCH((k_level_presmooth<1, 6>), blocks_for(l1.N, B), B, 0, l1.binv.p, l1.r.p, mg.om(1), l1.N, l1.z.p);
        if (mg.K == 1) {
            if (flow) MOF_LAUNCH(k_dense_restrict_apply<3>, DENSE_CTAS, B, 0, (const int*)nullptr, l1.r.p, mg.cinv.p, 3 * l1.N, l1.z.p);
            else MOF_LAUNCH(k_dense_restrict_apply<6>, DENSE_CTAS, B, 0, (const int*)nullptr, l1.r.p, mg.cinv.p, l1.N, l1.z.p);
        }
        MOF_TRY(coarse_cycle(ctx, mg, 0));
        }
        if (rows) {
            if (flow) MOF_LAUNCH(k_prolong_flow, blocks_for(rows, B), B, 0, mg.agg.p + r0, mg.cevec.p + 3 * (size_t)r0, l1.z.p, rows, mg.fz.p + e0);
            else MOF_LAUNCH(k_prolong_scalar, blocks_for(len, B), B, 0, mg.agg.p + r0, l1.z.p, rows, mg.fz.p + e0);
        }
        MOF_TRY(dist_halo(ctx, kind, mg.fz.p));
        MOF_TRY(spmv((const creal*)mg.fval.p, r, mg.om(0), (const creal*)mg.fz.p, mg.fz2.p, 2, partials));
        if (withRR) {
            MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, FINE_GRID, R_RZNEW, mg.scal.p);
            MOF_TRY(dist_allreduce(ctx, mg.scal.p + R_RR, 2));
            MOF_LAUNCH(k_derive, 1, 1, 0, R_RR, mg.scal.p);
            MOF_LAUNCH(k_derive, 1, 1, 0, R_RZNEW, mg.scal.p);
        } else
            MOF_TRY(reduce(FINE_GRID, rawSlot));
        std::swap(mg.fz.p, mg.fz2.p);
        return MOF_OK;
    };

    if (zeroGuess) {
        MOF_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * full, ctx->stream));
        MOF_CUDA(cudaMemcpyAsync(r, b, sizeof(double) * full, cudaMemcpyDeviceToDevice, ctx->stream));
    } else  // x is complete on every rank: no exchange needed for the first residual
        MOF_TRY(spmv(sysVal, b, mg.om(OM_ZERO), (const double*)x, r, 1, NO_FOLD));
    MOF_LAUNCH((k_dot_partial<double, double>), NBLK, B, 0, b + e0, b + e0, len, mg.partial.p);
    MOF_TRY(reduce(NBLK, R_BB));
    MOF_LAUNCH((k_dot_partial<double, double>), NBLK, B, 0, r + e0, r + e0, len, mg.partial.p);
    MOF_TRY(reduce(NBLK, R_RR));
    double h2[2] = {0, 0};
    MOF_CUDA(read_back(ctx, h2, mg.scal.p + S_RR, 2));
    double rr = h2[0];
    const double bb = h2[1];
    *itersOut = 0, *relresOut = 0;
    if (!(bb > 0)) {
        if (!zeroGuess) MOF_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * full, ctx->stream));
        return MOF_OK;
    }
    int it = 0;
    long long launchesPerReplay = 0;
    for (int attempt = 0; attempt < 4 && rr > tol * tol * bb; attempt++) {
        MOF_TRY(cycle(false, R_RZ));
        if (len) MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p + e0, mg.scal.p, len, 1, p + e0);
        auto iteration = [&](int half) -> int {
            MOF_TRY(dist_halo(ctx, kind, p));
            MOF_TRY(spmv(sysVal, (const double*)nullptr, mg.om(OM_ZERO), (const double*)p, q, 0, partials));
            MOF_TRY(reduce(FINE_GRID, R_PQ));
            MOF_LAUNCH(k_update_xr, NBLK, B, 0, p + e0, q + e0, len, x + e0, r + e0, mg.fdinv.p + r0, mg.om(0), W, mg.fz.p + e0, partials);
            MOF_LAUNCH(k_fold, 1, B, 0, mg.partial.p, NBLK, R_RR, mg.scal.p);  // this rank's part; all-reduced with r.z at the end of the cycle
            MOF_TRY(cycle(true, R_RZNEW, true));
            MOF_CUDA(cudaMemcpyAsync(mg.hostRR + half, mg.scal.p + S_RR, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            if (len) MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p + e0, mg.scal.p, len, 0, p + e0);
            return MOF_OK;
        };
        // Two iterations (kernels and NCCL operations alike) captured as one CUDA graph and replayed, as in mg_pcg;
        // MOF_DIST_GRAPH=0 launches them one by one.
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        if (env_int("MOF_DIST_GRAPH", 1)) {
            const long long launchesBefore = ctx->stats.kernelLaunches;
            MOF_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
            int crc = iteration(0);
            if (crc == MOF_OK) crc = iteration(1);
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
            launchesPerReplay = ctx->stats.kernelLaunches - launchesBefore;
            ctx->stats.kernelLaunches = launchesBefore;
            if (crc != MOF_OK || ce != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                return crc != MOF_OK ? crc : cuda_fail(ctx, ce, "cudaStreamEndCapture(partitioned iteration)");
            }
            ce = cudaGraphInstantiate(&exec, graph, 0);
            if (ce != cudaSuccess) {
                cudaGraphDestroy(graph);
                return cuda_fail(ctx, ce, "cudaGraphInstantiate(partitioned iteration)");
            }
        }
        int lrc = MOF_OK;
        while (it < maxIters) {
            if (exec) {
                cudaError_t ce = cudaGraphLaunch(exec, ctx->stream);
                if (ce != cudaSuccess) { lrc = cuda_fail(ctx, ce, "cudaGraphLaunch(partitioned iteration)"); break; }
                ctx->stats.kernelLaunches += launchesPerReplay;
            } else {
                lrc = iteration(0);
                if (lrc == MOF_OK) lrc = iteration(1);
                if (lrc != MOF_OK) break;
            }
            cudaError_t se = cudaStreamSynchronize(ctx->stream);
            if (se != cudaSuccess) { lrc = cuda_fail(ctx, se, "cudaStreamSynchronize(partitioned iteration)"); break; }
            it += 2;
            rr = mg.hostRR[1];
            if (!(mg.hostRR[0] > tol * tol * bb) || !(rr > tol * tol * bb) || !std::isfinite(rr)) break;
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        if (lrc != MOF_OK) return lrc;
        // true residual of x
        MOF_TRY(dist_halo(ctx, kind, x));
        MOF_TRY(spmv(sysVal, b, mg.om(OM_ZERO), (const double*)x, r, 1, NO_FOLD));
        MOF_LAUNCH((k_dot_partial<double, double>), NBLK, B, 0, r + e0, r + e0, len, mg.partial.p);
        MOF_TRY(reduce(NBLK, R_RR));
        MOF_CUDA(read_back(ctx, &rr, mg.scal.p + S_RR));
        if (it >= maxIters || !std::isfinite(rr)) break;
    }
    MOF_TRY(dist_allgather_rows(ctx, kind, x));
    MOF_TRY(dist_p2p_check(ctx));
    *itersOut = it;
    *relresOut = std::sqrt(rr / bb);
    if (!(*relresOut <= tol * 1.0001)) ctx->stats.solvesAboveTolerance++;
    if (!std::isfinite(rr) || (!(*relresOut <= tol * 1.0001) && (it >= maxIters || !(*relresOut <= MOF_ACCEPT_RELRES)))) {
        char msg[160];
        snprintf(msg, sizeof(msg), "[ERROR] partitioned multigrid PCG did not reach %g in %d iterations (relative residual %g)", tol, it, *relresOut);
        return fail(ctx, MOF_E_NOCONVERGE, msg);
    }
    return MOF_OK;
}

}  // namespace

int mg_dist_setup(mof_ctx* ctx) {
    if (!dist_active(ctx) || dist_world(ctx) < 2) {
        if (ctx->mg) ctx->mg->distLevels = 0;
        if (ctx->mgs) ctx->mgs->distLevels = 0;
        return MOF_OK;
    }
    if (ctx->mg) MOF_TRY(mg_dist_setup_one(ctx, *ctx->mg));
    if (ctx->mgs) MOF_TRY(mg_dist_setup_one(ctx, *ctx->mgs));
    return MOF_OK;
}

// FLOW: coarse operators of the current wA, then the solve wA x = fb into fx.
int mg_flow_update(mof_ctx* ctx) {
    Multigrid& mg = *ctx->mg;
    MgLevel& l1 = mg.lev[0];
    MOF_LAUNCH(k_level1_flow, blocks_for(27ll * l1.N, B), B, 0, mg.aggPtr.p, mg.aggList.p, ctx->wRowptr.p, ctx->wSliceBase.p, ctx->wCol.p, ctx->wA.p, mg.slotOf.p, mg.evec.p,
               l1.nbr.p, l1.N, coarse_scale(mg, 0), l1.blocks.p);
    MOF_LAUNCH(k_to_creal, kSMs * 8, B, 0, ctx->wA.p, ctx->wPadded, mg.fval.p);
    MOF_LAUNCH(k_to_creal, kSMs * 8, B, 0, ctx->wDinv.p, (long long)ctx->E, mg.fdinv.p);
    return finish_values(ctx, mg);
}
int mg_flow_solve(mof_ctx* ctx, double tol, int maxIters, int* itersOut, double* relresOut) {
    if (dist_active(ctx)) return mg_pcg_dist(ctx, *ctx->mg, ctx->fb.p, ctx->fx.p, true, tol, maxIters, itersOut, relresOut);
    return mg_pcg(ctx, *ctx->mg, ctx->fb.p, ctx->fx.p, true, tol, maxIters, itersOut, relresOut);
}

// SCALAR: coarse operators of the current sSys = M + eps S (with sDinv its inverse diagonal), then the six-channel
// solve sSys X = b6 starting from the content of x6.
int mg_scalar_update(mof_ctx* ctx) {
    Multigrid& mg = *ctx->mgs;
    MgLevel& l1 = mg.lev[0];
    MOF_LAUNCH(k_level1_scalar, blocks_for(27ll * l1.N, B), B, 0, mg.aggPtr.p, mg.aggList.p, ctx->sRowptr.p, ctx->sSys.p, mg.slotOf.p, l1.nbr.p, l1.N, coarse_scale(mg, 0), l1.blocks.p);
    MOF_LAUNCH(k_to_creal, kSMs * 8, B, 0, ctx->sSys.p, ctx->nnzS, mg.fval.p);
    MOF_LAUNCH(k_to_creal, kSMs * 8, B, 0, ctx->sDinv.p, (long long)ctx->V, mg.fdinv.p);
    if (ctx->sPadded > 0) {
        const int slices = (ctx->V + 31) / 32;
        MOF_LAUNCH(k_scalar_vals_to_sell, blocks_for(32ll * slices, B), B, 0, ctx->sRowptr.p, ctx->sSys.p, ctx->sSliceBase.p, ctx->V, slices, ctx->sSysSell.p, mg.fvalSell.p);
    }
    return finish_values(ctx, mg);
}
// One cycle of the SCALAR hierarchy as an approximate solve of sSys Z = R from a zero guess (six channels, [V][6]): a fixed,
// symmetric positive definite operator (to the fp32 of the cycle). The coarse operators must be those of the current sSys
// (mg_scalar_update). A building block for preconditioners outside this file (the Conformal basis, vector_fields.cu).
int mg_scalar_cycle(mof_ctx* ctx, const double* r6, double* z6) {
    Multigrid& mg = *ctx->mgs;
    MOF_TRY(fine_cycle(ctx, mg, r6, false, S_RZ));
    const long long len = (long long)mg.fineLen();
    MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p, mg.scal.p, len, 1, z6);
    return MOF_OK;
}
// mg_flow_update for a matrix that is not an alignment's (spectrum.cu): whether the hierarchy took it (positive definite coarsest level);
// either way the hierarchy stays available to the next alignment, whose own system re-values it.
bool mg_flow_try_update(mof_ctx* ctx) {
    if (!mg_flow_usable(ctx)) return false;
    const int rc = mg_flow_update(ctx);
    const bool ok = rc == MOF_OK && ctx->mg->usable;
    ctx->mg->usable = true;
    return ok;
}
// The same for the FLOW hierarchy: Z = cycle(R) for the matrix in ctx->wA whose coarse operators mg_flow_update built last — the
// preconditioner of the Spectrum tool's block iteration (spectrum.cu), where wA = S + tau M.
int mg_flow_cycle(mof_ctx* ctx, const double* r, double* z) {
    Multigrid& mg = *ctx->mg;
    MOF_TRY(fine_cycle(ctx, mg, r, false, S_RZ));
    const long long len = (long long)mg.fineLen();
    MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p, mg.scal.p, len, 1, z);
    return MOF_OK;
}
// d = alpha d + beta c ; z += d (one step of the Chebyshev recurrence below)
__global__ void k_cheb_update(double alpha, double beta, const double* __restrict__ c, long long n, double* __restrict__ d, double* __restrict__ z, int first) { pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dn = first ? beta * c[i] : alpha * d[i] + beta * c[i];
    d[i] = dn;
    z[i] = first ? dn : z[i] + dn;
}
// A sharper approximate inverse of the current sSys than one cycle: `degree` steps of the Chebyshev iteration for sSys Z = R
// preconditioned by the cycle C (degree 1 = one cycle, scaled). The eigenvalues of C sSys lie in (0, 1] for a symmetric cycle
// with convergent smoothers (its error propagator is sSys-self-adjoint and non-negative); [lo, 1.05] is what the polynomial is
// built for — below lo it merely reduces less. Fixed coefficients, no inner products: the result is a FIXED symmetric positive
// definite operator applied to R, so plain PCG may use it as (part of) a preconditioner. Used by the Conformal basis, whose
// bi-Laplacian preconditioner squares the inverse's error (vector_fields.cu).
int mg_scalar_cheb(mof_ctx* ctx, const double* r6, double* z6, int degree, double lo) {
    Multigrid& mg = *ctx->mgs;
    const long long len = (long long)mg.fineLen();
    if (degree <= 1) return mg_scalar_cycle(ctx, r6, z6);
    degree = std::min(degree, 32);
    MOF_CUDA(mg.chebD.reserve((size_t)len));
    MOF_CUDA(mg.chebR.reserve((size_t)len));
    MOF_CUDA(mg.chebC.reserve((size_t)len));
    const double hi = 1.05, theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma1 = theta / delta;
    double rho = 1. / sigma1;
    MOF_TRY(mg_scalar_cycle(ctx, r6, mg.chebC.p));
    MOF_LAUNCH(k_cheb_update, blocks_for(len, B), B, 0, 0., 1. / theta, mg.chebC.p, len, mg.chebD.p, z6, 1);
    for (int k = 1; k < degree; k++) {
        MOF_TRY(fine_residual(ctx, mg, r6, z6, mg.chebR.p));
        MOF_TRY(mg_scalar_cycle(ctx, mg.chebR.p, mg.chebC.p));
        const double rhoNew = 1. / (2. * sigma1 - rho);
        MOF_LAUNCH(k_cheb_update, blocks_for(len, B), B, 0, rhoNew * rho, 2. * rhoNew / delta, mg.chebC.p, len, mg.chebD.p, z6, 0);
        rho = rhoNew;
    }
    return MOF_OK;
}

// How good an inverse of the current sSys one cycle C is: the smallest eigenvalue of C sSys (the largest is <= 1 for a symmetric cycle
// with convergent smoothers), from the Lanczos tridiagonal that `steps` iterations of cycle-preconditioned CG on a pseudo-random
// right-hand side carry in their alpha / beta (diagonal 1/alpha_j + beta_{j-1}/alpha_{j-1}, off-diagonal sqrt(beta_j)/alpha_j): its
// smallest Ritz value approaches the smallest eigenvalue from above, and — unlike a power iteration on I - C sSys, which returned 0.917
// for a cycle that really contracts by 0.99 at 1M vertices — does so quickly at the ends of the spectrum.
__global__ void k_pseudo_random_f64(long long n, double* __restrict__ v) { pdl_wait();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = (unsigned)i * 2654435761u + 12345u;
    h ^= h >> 15, h *= 2246822519u, h ^= h >> 13;
    v[i] = (double)(h & 0xffff) / 32768. - 1.;
}
int mg_scalar_smallest_eigenvalue(mof_ctx* ctx, int steps, double* lambdaMin) {
    Multigrid& mg = *ctx->mgs;
    const long long len = (long long)mg.fineLen();
    MOF_CUDA(mg.chebD.reserve((size_t)len));
    double* x = mg.chebD.p;  // (the iterate itself is of no interest)
    double* r = mg.fr.p;
    double* p = mg.fp.p;
    double* q = mg.fq.p;
    MOF_LAUNCH(k_pseudo_random_f64, blocks_for(len, B), B, 0, len, r);
    MOF_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * len, ctx->stream));
    MOF_TRY(fine_cycle(ctx, mg, r, false, S_RZ));
    MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p, mg.scal.p, len, 1, p);
    std::vector<double> alpha, beta;
    for (int it = 0; it < steps; it++) {
        MOF_TRY(apply_dot(ctx, mg, p, q));
        MOF_LAUNCH(k_update_xr, NBLK, B, 0, p, q, len, x, r, mg.fdinv.p, mg.om(0), mg.nrhs, mg.fz.p, fold_into(mg, S_RR));
        MOF_TRY(fold_after(ctx, mg, NBLK, S_RR));
        MOF_TRY(fine_cycle(ctx, mg, r, true, S_RZNEW));
        MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p, mg.scal.p, len, 0, p);
        double ab[2] = {0, 0};
        MOF_CUDA(read_back(ctx, ab, mg.scal.p + S_ALPHA, 2));  // S_ALPHA, S_BETA are adjacent
        if (!(ab[0] > 0) || !(ab[1] >= 0) || !std::isfinite(ab[0]) || !std::isfinite(ab[1])) break;
        alpha.push_back(ab[0]), beta.push_back(ab[1]);
    }
    const int m = (int)alpha.size();
    if (m < 2) { *lambdaMin = 0.1; return MOF_OK; }
    std::vector<double> d(m), e(m, 0.);
    for (int j = 0; j < m; j++) {
        d[j] = 1. / alpha[j] + (j ? beta[j - 1] / alpha[j - 1] : 0.);
        if (j + 1 < m) e[j] = std::sqrt(beta[j]) / alpha[j];
    }
    // smallest eigenvalue of the tridiagonal (d, e) by bisection on the Sturm count
    double lo = 0., hi = d[0];
    for (int j = 0; j < m; j++) hi = std::min(hi, d[j]);  // the smallest diagonal entry bounds the smallest eigenvalue from above
    for (int round = 0; round < 60; round++) {
        const double mid = 0.5 * (lo + hi);
        int below = 0;
        double piv = d[0] - mid;
        if (piv < 0) below++;
        for (int j = 1; j < m; j++) {
            piv = d[j] - mid - e[j - 1] * e[j - 1] / (piv != 0 ? piv : 1e-300);
            if (piv < 0) below++;
        }
        if (below >= 1) hi = mid; else lo = mid;
    }
    *lambdaMin = hi;
    return MOF_OK;
}

// mof_time_kernel: one kernel of a PCG iteration, launched `reps` times on its own with the solver's grid on the solver's own
// buffers (scratch: fr, fp, fq, fz, ft are re-initialised by every solve). *bytes = the algorithmic bytes of one launch.
int mg_time_kernel(mof_ctx* ctx, int which, int reps, float* ms, double* bytes) {
    const bool scalar = which >= MOF_K_SCALAR_SPMV && which <= MOF_K_SCALAR_LEVEL1;
    Multigrid* mgp = scalar ? ctx->mgs : ctx->mg;
    if (!mgp || !mgp->usable || mgp->K < 1) return fail(ctx, MOF_E_UNSUPPORTED, "mof_time_kernel: this hierarchy is not in use on this mesh");
    Multigrid& mg = *mgp;
    const long long len = (long long)mg.fineLen();
    const long long n = mg.nFine;
    MgLevel& l1 = mg.lev[0];
    const double nnz = scalar ? (double)ctx->nnzS : (double)ctx->wPadded;
    double* scratchX = ctx->pcg.q.p;
    auto launch = [&]() -> int {
        switch (which) {
            case MOF_K_FLOW_SPMV:
            case MOF_K_SCALAR_SPMV: return apply_dot(ctx, mg, mg.fp.p, mg.fq.p);
            case MOF_K_FLOW_FINE_SWEEP:
            case MOF_K_SCALAR_FINE_SWEEP: return fine_apply(ctx, mg, mg.fr.p, mg.om(0), mg.fz.p, mg.fz2.p, 2, S_RZNEW);
            case MOF_K_FLOW_UPDATE:
            case MOF_K_SCALAR_UPDATE:
                MOF_LAUNCH(k_update_xr, NBLK, B, 0, mg.fp.p, mg.fq.p, len, scratchX, mg.fr.p, mg.fdinv.p, mg.om(0), mg.nrhs, mg.fz.p, fold_into(mg, S_RR));
                return MOF_OK;
            case MOF_K_FLOW_RESTRICT:
                MOF_LAUNCH(k_restrict_flow, blocks_for(32ll * l1.N, B), B, 0, mg.aggPtr.p, mg.aggList.p, mg.cevecAgg.p, mg.ft.p, l1.N, l1.binv.p, mg.om(1), l1.r.p, l1.z.p, 0, mg.nFine, 0, -1);
                return MOF_OK;
            case MOF_K_FLOW_PROLONG:
                MOF_LAUNCH(k_prolong_flow, blocks_for(mg.nFine, B), B, 0, mg.agg.p, mg.cevec.p, l1.z.p, mg.nFine, mg.fz.p);
                return MOF_OK;
            case MOF_K_FLOW_DIRECTION:
                MOF_LAUNCH(k_direction, blocks_for(len, B), B, 0, mg.fz.p, mg.scal.p, len, 0, mg.fp.p);
                return MOF_OK;
            case MOF_K_FLOW_LEVEL1:
            case MOF_K_SCALAR_LEVEL1: return coarse_apply(ctx, mg, l1, mg.om(1), 2, l1.t.p);
        }
        return fail(ctx, MOF_E_INVALID, "mof_time_kernel: unknown kernel");
    };
    MOF_CUDA(ctx->pcg.q.reserve((size_t)len));
    scratchX = ctx->pcg.q.p;
    // defined inputs (zeros are as good as anything for timing; no NaN patterns from stale scratch)
    MOF_CUDA(cudaMemsetAsync(mg.fp.p, 0, sizeof(double) * len, ctx->stream));
    MOF_CUDA(cudaMemsetAsync(mg.fq.p, 0, sizeof(double) * len, ctx->stream));
    MOF_CUDA(cudaMemsetAsync(mg.fr.p, 0, sizeof(double) * len, ctx->stream));
    MOF_CUDA(cudaMemsetAsync(scratchX, 0, sizeof(double) * len, ctx->stream));
    MOF_CUDA(cudaMemsetAsync(mg.fz.p, 0, sizeof(creal) * len, ctx->stream));
    MOF_CUDA(cudaMemsetAsync(mg.ft.p, 0, sizeof(creal) * len, ctx->stream));
    MOF_CUDA(cudaMemsetAsync(l1.r.p, 0, sizeof(creal) * mg.dofs() * l1.N, ctx->stream));
    MOF_CUDA(cudaMemsetAsync(l1.z.p, 0, sizeof(creal) * mg.dofs() * l1.N, ctx->stream));
    MOF_CUDA(cudaMemsetAsync(mg.scal.p, 0, sizeof(double) * 8, ctx->stream));
    for (int i = 0; i < 3; i++) MOF_TRY(launch());
    MOF_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < reps; i++) MOF_TRY(launch());
    MOF_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    MOF_CUDA(cudaEventSynchronize(ctx->ev1));
    MOF_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    *ms /= reps;
    const double c = sizeof(creal), N1 = l1.N;
    switch (which) {
        case MOF_K_FLOW_SPMV: *bytes = 12. * ctx->nnzW + 4. * (n + 1) + 16. * n; break;
        case MOF_K_FLOW_FINE_SWEEP: *bytes = (c + 4.) * nnz + 4. * (ctx->wSlices + 1) + n * (8. + 3. * c); break;  // values + columns; b, dinv, in, out
        case MOF_K_FLOW_UPDATE:
        case MOF_K_SCALAR_UPDATE: *bytes = len * (6. * 8. + c) + n * c; break;                                  // p, q, x, r in; x, r, z out; dinv
        case MOF_K_FLOW_RESTRICT: *bytes = n * (4. + c + 3. * c) + N1 * (9. * c + 6. * c); break;
        case MOF_K_FLOW_PROLONG: *bytes = n * (4. + 3. * c + 2. * c) + N1 * 3. * c; break;
        case MOF_K_FLOW_DIRECTION: *bytes = len * (c + 16.); break;
        case MOF_K_FLOW_LEVEL1: *bytes = N1 * (27. * 9. * c + 27. * 4. + 9. * c + 3. * 3. * c); break;
        case MOF_K_SCALAR_LEVEL1: *bytes = N1 * (27. * c + 27. * 4. + c + 3. * 6. * c); break;
        case MOF_K_SCALAR_SPMV: *bytes = 12. * nnz + 4. * (n + 1) + 6. * 16. * n; break;
        case MOF_K_SCALAR_FINE_SWEEP: *bytes = (c + 4.) * nnz + 4. * (n + 1) + n * (6. * 8. + c + 2. * 6. * c); break;
        default: *bytes = 0;
    }
    return MOF_OK;
}
int mg_scalar_solve(mof_ctx* ctx, const double* b6, double* x6, double tol, int maxIters, int* itersOut, double* relresOut) {
    if (dist_active(ctx)) return mg_pcg_dist(ctx, *ctx->mgs, b6, x6, false, tol, maxIters, itersOut, relresOut);
    return mg_pcg(ctx, *ctx->mgs, b6, x6, false, tol, maxIters, itersOut, relresOut);
}

}  // namespace mof
