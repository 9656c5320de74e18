// Device code of vector_fields.cu (the Conformal and Connection bases), kept free of the CUDA runtime API so that
// tests/test_vf_kernels_host.py can compile the element-wise kernels for the HOST (MOF_HOST_EMULATION: __global__
// and friends defined away, one "thread" at a time) and check this very source against the CPU checker without a
// GPU. The reduction kernels (shared memory, barriers) are not emulated.
#pragma once

#include "mof_slots.h"

#ifndef MOF_HOST_EMULATION
#include <cuda_runtime.h>
#endif

namespace mof {
namespace vfk {

constexpr int B = 256;
#ifdef MOF_HOST_EMULATION
constexpr int RED = 3;
#else
constexpr int RED = kSMs * 4;  // partial sums per reduction
#endif

enum { S_RZ0 = 0, S_RZ1 = 1, S_PQ = 2, S_RR = 3, S_BB = 4, S_COUNT = 8 };

__device__ __forceinline__ double det3(const double* g) { return g[0] * g[2] - g[1] * g[1]; }
__device__ __forceinline__ void inv3(const double* g, double* gi) {
    double d = 1. / det3(g);
    gi[0] = g[2] * d, gi[1] = -g[1] * d, gi[2] = g[0] * d;
}
__device__ __forceinline__ double gdot(const double* g, double ax, double ay, double bx, double by) {
    return ax * (g[0] * bx + g[1] * by) + ay * (g[1] * bx + g[2] * by);
}

// Row k of the Conformal prolongation restricted to triangle t (Conformal.inl:55-79): the flow contributed by
// a unit potential (pg) and a unit co-potential (pr) at corner k.
struct ConformalP {
    double pg[3][2], pr[3][2];
};
__device__ __forceinline__ void conformal_p(const double* __restrict__ g, int t, ConformalP& p) {
    const double gx[3] = {-1., 1., 0.}, gy[3] = {-1., 0., 1.};      // hat gradients
    const double rx[3] = {1., 0., -1.}, ry[3] = {-1., 1., 0.};      // rotated gradients (Conformal.inl:56)
    double gt[3] = {g[3 * t], g[3 * t + 1], g[3 * t + 2]}, gi[3];
    inv3(gt, gi);
    double is = 1.0 / sqrt(det3(gt));
#pragma unroll
    for (int k = 0; k < 3; k++) {
        p.pg[k][0] = gi[0] * gx[k] + gi[1] * gy[k], p.pg[k][1] = gi[1] * gx[k] + gi[2] * gy[k];
        p.pr[k][0] = rx[k] * is, p.pr[k][1] = ry[k] * is;
    }
}
// One corner only (the per-vertex gathers).
__device__ __forceinline__ void conformal_corner(const double* __restrict__ g, int t, int k, double* pg, double* pr) {
    const double gx[3] = {-1., 1., 0.}, gy[3] = {-1., 0., 1.};
    const double rx[3] = {1., 0., -1.}, ry[3] = {-1., 1., 0.};
    double gt[3] = {g[3 * t], g[3 * t + 1], g[3 * t + 2]}, gi[3];
    inv3(gt, gi);
    double is = 1.0 / sqrt(det3(gt));
    pg[0] = gi[0] * gx[k] + gi[1] * gy[k], pg[1] = gi[1] * gx[k] + gi[2] * gy[k];
    pr[0] = rx[k] * is, pr[1] = ry[k] * is;
}
__device__ __forceinline__ double quad(const double* D, const double* a, const double* b) {  // a^T D b, D = (d00,d01,d11)
    return a[0] * (D[0] * b[0] + D[1] * b[1]) + a[1] * (D[1] * b[0] + D[2] * b[1]);
}

// ------------------------------------------------------------------------------------ reductions

__global__ void k_dot_partial(const double* __restrict__ a, const double* __restrict__ b, long long n, double* __restrict__ partial) {
    __shared__ double sh[B];
    double s = 0;
    for (long long i = (long long)blockIdx.x * B + threadIdx.x; i < n; i += (long long)gridDim.x * B) s += a[i] * (b ? b[i] : 1.0);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (!threadIdx.x) partial[blockIdx.x] = sh[0];
}
// out[j] = sum_i partial[j * np + i], j < count: one block, fixed order.
__global__ void k_fold(const double* __restrict__ partial, int np, int count, double* __restrict__ out0, double* __restrict__ out1) {
    __shared__ double sh[B];
    for (int j = 0; j < count; j++) {
        double s = 0;
        for (int i = threadIdx.x; i < np; i += B) s += partial[(size_t)j * np + i];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int o = B / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (!threadIdx.x) *(j ? out1 : out0) = sh[0];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ Connection: set-up

// Connection.inl:41-99 for one triangle: the edge weights l_j, the diagonal block sum_j l_j g and the three
// transport blocks -l_j g L_j (L_j = linear part of the transform from the neighbour across edge j into this chart).
__global__ void k_connection_blocks(const double* __restrict__ g, const double* __restrict__ area, const int* __restrict__ opp, const double* __restrict__ xlin,
                                    const double* __restrict__ xcst, int cMode, int T, double* __restrict__ diag, double* __restrict__ off) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const double ex[3] = {-1., 0., 1.}, ey[3] = {1., -1., 0.};  // Connection.inl:35
    double gt[3] = {g[3 * t], g[3 * t + 1], g[3 * t + 2]};
    double at = area[t];
    double d0 = 0, d1 = 0, d2 = 0;
    for (int j = 0; j < 3; j++) {
        int o = opp[3 * t + j], ii = o / 3, jj = o - 3 * ii;
        const double* L = xlin + 4 * (size_t)o;
        double aii = area[ii], l;
        if (cMode == 0)
            l = gdot(gt, ex[j], ey[j], ex[j], ey[j]) / (4.0 * (at + aii) / 3.0);
        else if (cMode == 1) {
            const double c = 1. / 3;
            double dx = c - (L[0] * c + L[1] * c + xcst[2 * (size_t)o]), dy = c - (L[2] * c + L[3] * c + xcst[2 * (size_t)o + 1]);
            l = ((at + aii) / 3.0) / gdot(gt, dx, dy, dx, dy);
        } else {
            double g2[3] = {g[3 * ii], g[3 * ii + 1], g[3 * ii + 2]};
            int j1 = (j + 1) % 3, j2 = (j + 2) % 3, k1 = (jj + 1) % 3, k2 = (jj + 2) % 3;
            l = 1.0 / (gdot(gt, -ex[j1], -ey[j1], ex[j2], ey[j2]) / (2.0 * at) + gdot(g2, -ex[k1], -ey[k1], ex[k2], ey[k2]) / (2.0 * aii));
        }
        d0 += l * gt[0], d1 += l * gt[1], d2 += l * gt[2];
        double* X = off + 12 * (size_t)t + 4 * j;
        X[0] = -l * (gt[0] * L[0] + gt[1] * L[2]), X[1] = -l * (gt[0] * L[1] + gt[1] * L[3]);
        X[2] = -l * (gt[1] * L[0] + gt[2] * L[2]), X[3] = -l * (gt[1] * L[1] + gt[2] * L[3]);
    }
    diag[3 * t] = d0, diag[3 * t + 1] = d1, diag[3 * t + 2] = d2;
}

// ---------------------------------------------------------------------- per-iteration system pieces

// Connection: R D P = D (block diagonal), so ||.||_F^2 = sum d00^2 + 2 d01^2 + d11^2 (VectorField.h:57).
__global__ void k_connection_frob(const double* __restrict__ D, int T, double* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    out[t] = D[3 * t] * D[3 * t] + 2. * D[3 * t + 1] * D[3 * t + 1] + D[3 * t + 2] * D[3 * t + 2];
}

// Conformal, row pair (a_v, b_v) of R D P = P^T D P and of the bi-Laplacian diagonal:
//   rowSq[v]   sum of squares of the merged entries of both rows (every neighbour u: the 2x2 block summed over the
//              two triangles on edge vu; the diagonal block summed over the fan), for the Frobenius norm;
//   blk[v][4]  unscaled diagonal block (aa, ab, bb) of P^T D P and the bi-Laplacian diagonal sum_k K_vk^2 / m_k.
__global__ void k_conformal_rows(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ he, const int* __restrict__ opp,
                                 const double* __restrict__ stiff, const double* __restrict__ minv, const double* __restrict__ g, const double* __restrict__ D, int V,
                                 double* __restrict__ rowSq, double* __restrict__ blk) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double sq = 0, daa = 0, dab = 0, dbb = 0, bil = 0;
    for (int k = rowptr[v]; k < rowptr[v + 1]; k++) {
        double kv = stiff[k];
        bil += kv * kv * minv[col[k]];
        int h = he[k];
        if (h < 0) continue;
        // triangle t: v is corner j+1, the neighbour u corner j+2
        int t = h / 3, j = h - 3 * t;
        double pgv[2], prv[2], pgu[2], pru[2];
        conformal_corner(g, t, (j + 1) % 3, pgv, prv);
        conformal_corner(g, t, (j + 2) % 3, pgu, pru);
        const double* Dt = D + 3 * (size_t)t;
        double eaa = quad(Dt, pgv, pgu), eab = quad(Dt, pgv, pru), eba = quad(Dt, prv, pgu), ebb = quad(Dt, prv, pru);
        daa += quad(Dt, pgv, pgv), dab += quad(Dt, pgv, prv), dbb += quad(Dt, prv, prv);
        int o = opp[h];
        if (o >= 0) {  // triangle t2 across the edge: u is corner j2+1, v corner j2+2
            int t2 = o / 3, j2 = o - 3 * t2;
            conformal_corner(g, t2, (j2 + 2) % 3, pgv, prv);
            conformal_corner(g, t2, (j2 + 1) % 3, pgu, pru);
            const double* D2 = D + 3 * (size_t)t2;
            eaa += quad(D2, pgv, pgu), eab += quad(D2, pgv, pru), eba += quad(D2, prv, pgu), ebb += quad(D2, prv, pru);
        }
        sq += eaa * eaa + eab * eab + eba * eba + ebb * ebb;
    }
    sq += daa * daa + 2. * dab * dab + dbb * dbb;
    rowSq[v] = sq;
    blk[4 * (size_t)v] = daa, blk[4 * (size_t)v + 1] = dab, blk[4 * (size_t)v + 2] = dbb, blk[4 * (size_t)v + 3] = bil;
}

__global__ void k_set_scale(double* __restrict__ scalars) { scalars[SC_DATA_SCALE] = 1. / sqrt(scalars[SC_FROB2]); }

__device__ __forceinline__ void invert_block(double a, double b, double c, double* out) {
    double d = 1. / (a * c - b * b);
    out[0] = c * d, out[1] = -b * d, out[2] = a * d;
}

// Block-Jacobi inverses and the scaled right-hand side s * R rhs (VectorField.h:53, 60).
__global__ void k_connection_finalize(const double* __restrict__ D, const double* __restrict__ rhs, const double* __restrict__ diag, const double* __restrict__ scalars,
                                      double weight, int T, double* __restrict__ binv, double* __restrict__ b) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double s = scalars[SC_DATA_SCALE];
    invert_block(s * D[3 * t] + weight * diag[3 * t], s * D[3 * t + 1] + weight * diag[3 * t + 1], s * D[3 * t + 2] + weight * diag[3 * t + 2], binv + 3 * (size_t)t);
    b[2 * t] = s * rhs[2 * t], b[2 * t + 1] = s * rhs[2 * t + 1];
}
__global__ void k_conformal_finalize(const int* __restrict__ rowptr, const int* __restrict__ he, const double* __restrict__ g, const double* __restrict__ rhs,
                                     const double* __restrict__ blk, const double* __restrict__ scalars, double weight, int V, double* __restrict__ binv,
                                     double* __restrict__ b) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double s = scalars[SC_DATA_SCALE];
    double ba = 0, bb = 0;
    for (int k = rowptr[v]; k < rowptr[v + 1]; k++) {
        int h = he[k];
        if (h < 0) continue;
        int t = h / 3, j = h - 3 * t;
        double pg[2], pr[2];
        conformal_corner(g, t, (j + 1) % 3, pg, pr);
        ba += pg[0] * rhs[2 * t] + pg[1] * rhs[2 * t + 1], bb += pr[0] * rhs[2 * t] + pr[1] * rhs[2 * t + 1];
    }
    b[v] = s * ba, b[v + V] = s * bb;
    const double* q = blk + 4 * (size_t)v;
    double bil = 0.5 * weight * q[3];
    invert_block(s * q[0] + bil, s * q[1], s * q[2] + bil, binv + 3 * (size_t)v);
}

// ------------------------------------------------------------------------------ operator application
//
// The apply kernels run on a FIXED grid of RED CTAs (grid-stride loops) and leave the partial sums of x . (A x) in
// partial[0..RED): the PCG's p.q costs no launch of its own, and its consumer folds the RED partials itself.

// Sum over the CTA, returned to every thread (fixed order: deterministic).
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[B];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = B / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    double s = sh[0];
    __syncthreads();
    return s;
}
// Sum of np partials; every CTA that calls it adds them in the same order and gets the same bits.
__device__ __forceinline__ double fold_partials(const double* __restrict__ partial, int np) {
    double s = 0;
    for (int i = threadIdx.x; i < np; i += B) s += partial[i];
    return block_sum(s);
}

// Connection: y_t = (s D_t + w Sdiag_t) x_t + w sum_j Soff_tj x_nbr(j).
__global__ void k_connection_apply(const double* __restrict__ D, const double* __restrict__ diag, const double* __restrict__ off, const int* __restrict__ opp,
                                   const double* __restrict__ scalars, double weight, const double* __restrict__ x, int T, double* __restrict__ y,
                                   double* __restrict__ partial) {
    const double s = scalars[SC_DATA_SCALE];
    double dot = 0;
    for (int t = blockIdx.x * B + threadIdx.x; t < T; t += gridDim.x * B) {
        double a00 = s * D[3 * t] + weight * diag[3 * t], a01 = s * D[3 * t + 1] + weight * diag[3 * t + 1], a11 = s * D[3 * t + 2] + weight * diag[3 * t + 2];
        double x0 = x[2 * t], x1 = x[2 * t + 1];
        double y0 = a00 * x0 + a01 * x1, y1 = a01 * x0 + a11 * x1;
        double o0 = 0, o1 = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            int n = opp[3 * t + j] / 3;
            const double* X = off + 12 * (size_t)t + 4 * j;
            double n0 = x[2 * n], n1 = x[2 * n + 1];
            o0 += X[0] * n0 + X[1] * n1, o1 += X[2] * n0 + X[3] * n1;
        }
        y0 += weight * o0, y1 += weight * o1;
        y[2 * t] = y0, y[2 * t + 1] = y1;
        dot += x0 * y0 + x1 * y1;
    }
    dot = block_sum(dot);
    if (!threadIdx.x) partial[blockIdx.x] = dot;
}

// Conformal, stage 1, one launch for two independent loops: triangles w_t = s D_t (P x)_t, vertices u = diag(1/m) K x
// on both halves.
__global__ void k_conformal_stage1(const int* __restrict__ tri, const double* __restrict__ g, const double* __restrict__ D, const double* __restrict__ scalars,
                                   const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ stiff, const double* __restrict__ minv,
                                   const double* __restrict__ x, int V, int T, double* __restrict__ w, double* __restrict__ u) {
    const double s = scalars[SC_DATA_SCALE];
    for (int t = blockIdx.x * B + threadIdx.x; t < T; t += gridDim.x * B) {
        ConformalP p;
        conformal_p(g, t, p);
        double z[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int v = tri[3 * t + k];
            double a = x[v], b = x[v + V];
            z[0] += p.pg[k][0] * a + p.pr[k][0] * b, z[1] += p.pg[k][1] * a + p.pr[k][1] * b;
        }
        const double* Dt = D + 3 * (size_t)t;
        w[2 * t] = s * (Dt[0] * z[0] + Dt[1] * z[1]), w[2 * t + 1] = s * (Dt[1] * z[0] + Dt[2] * z[1]);
    }
    for (int v = blockIdx.x * B + threadIdx.x; v < V; v += gridDim.x * B) {
        double a = 0, b = 0;
        for (int k = rowptr[v]; k < rowptr[v + 1]; k++) {
            double kv = stiff[k];
            int c = col[k];
            a += kv * x[c], b += kv * x[c + V];
        }
        u[2 * (size_t)v] = a * minv[v], u[2 * (size_t)v + 1] = b * minv[v];
    }
}
// Conformal, stage 2 (vertices): y = w/2 K u + P^T w, with the partials of x . y.
__global__ void k_conformal_row(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ he, const double* __restrict__ stiff,
                                const double* __restrict__ g, const double* __restrict__ w, const double* __restrict__ u, double weight, const double* __restrict__ x,
                                int V, double* __restrict__ y, double* __restrict__ partial) {
    double dot = 0;
    for (int v = blockIdx.x * B + threadIdx.x; v < V; v += gridDim.x * B) {
        double ka = 0, kb = 0, ya = 0, yb = 0;
        for (int k = rowptr[v]; k < rowptr[v + 1]; k++) {
            double kv = stiff[k];
            int c = col[k];
            ka += kv * u[2 * (size_t)c], kb += kv * u[2 * (size_t)c + 1];
            int h = he[k];
            if (h < 0) continue;
            int t = h / 3, j = h - 3 * t;
            double pg[2], pr[2];
            conformal_corner(g, t, (j + 1) % 3, pg, pr);
            ya += pg[0] * w[2 * t] + pg[1] * w[2 * t + 1], yb += pr[0] * w[2 * t] + pr[1] * w[2 * t + 1];
        }
        ya += 0.5 * weight * ka, yb += 0.5 * weight * kb;
        y[v] = ya, y[v + V] = yb;
        dot += x[v] * ya + x[v + V] * yb;
    }
    dot = block_sum(dot);
    if (!threadIdx.x) partial[blockIdx.x] = dot;
}

// ------------------------------------------------------------------------------------------ PCG
//
// One iteration = apply (above; p.q partials) -> k_pcg_step -> k_pcg_direction, all on RED CTAs. Scalars live in
// sc[]: r.z alternates between two slots so that no launch reads a slot another CTA of the same launch writes.

// pair i of the unknowns: (2i, 2i+1) interleaved (Connection) or (i, i+half) split (Conformal)
__device__ __forceinline__ void pair_index(long long i, long long half, int split, long long& i0, long long& i1) {
    if (split) i0 = i, i1 = i + half;
    else i0 = 2 * i, i1 = 2 * i + 1;
}

// z = Binv r, partial sums of r.z (out[0..RED)) and r.r (out[RED..2RED)) (start of a solve or a restart).
__global__ void k_pcg_start(const double* __restrict__ binv, const double* __restrict__ r, long long half, int split, double* __restrict__ z,
                            double* __restrict__ out) {
    double srz = 0, srr = 0;
    for (long long i = (long long)blockIdx.x * B + threadIdx.x; i < half; i += (long long)gridDim.x * B) {
        long long i0, i1;
        pair_index(i, half, split, i0, i1);
        double r0 = r[i0], r1 = r[i1];
        const double* q = binv + 3 * i;
        double z0 = q[0] * r0 + q[1] * r1, z1 = q[1] * r0 + q[2] * r1;
        z[i0] = z0, z[i1] = z1;
        srz += r0 * z0 + r1 * z1, srr += r0 * r0 + r1 * r1;
    }
    srz = block_sum(srz), srr = block_sum(srr);
    if (!threadIdx.x) out[blockIdx.x] = srz, out[gridDim.x + blockIdx.x] = srr;
}
// alpha = rz / (sum of the p.q partials); x += alpha p; r -= alpha q; z = Binv r; partial sums of r.z and r.r.
__global__ void k_pcg_step(const double* __restrict__ binv, const double* __restrict__ sc, int rzSlot, const double* __restrict__ pqPartial, int np,
                           const double* __restrict__ p, const double* __restrict__ q, long long half, int split, double* __restrict__ x, double* __restrict__ r,
                           double* __restrict__ z, double* __restrict__ out) {
    const double pq = fold_partials(pqPartial, np);
    const double alpha = pq != 0 ? sc[rzSlot] / pq : 0.0;
    double srz = 0, srr = 0;
    for (long long i = (long long)blockIdx.x * B + threadIdx.x; i < half; i += (long long)gridDim.x * B) {
        long long i0, i1;
        pair_index(i, half, split, i0, i1);
        x[i0] += alpha * p[i0], x[i1] += alpha * p[i1];
        double r0 = r[i0] - alpha * q[i0], r1 = r[i1] - alpha * q[i1];
        r[i0] = r0, r[i1] = r1;
        if (binv) {  // block-Jacobi; with another preconditioner the caller computes z and the r.z partials afterwards
            const double* bq = binv + 3 * i;
            double z0 = bq[0] * r0 + bq[1] * r1, z1 = bq[1] * r0 + bq[2] * r1;
            z[i0] = z0, z[i1] = z1;
            srz += r0 * z0 + r1 * z1;
        }
        srr += r0 * r0 + r1 * r1;
    }
    srz = block_sum(srz), srr = block_sum(srr);
    if (!threadIdx.x) out[blockIdx.x] = srz, out[gridDim.x + blockIdx.x] = srr;
}
// Folds the r.z / r.r partials (CTA 0 publishes them in sc[newSlot], sc[S_RR]); p = z + (rzNew / rzOld) p, or p = z when
// oldSlot < 0.
__global__ void k_pcg_direction(double* __restrict__ sc, int newSlot, int oldSlot, const double* __restrict__ rzrr, int np, const double* __restrict__ z,
                                long long n, double* __restrict__ p) {
    const double rzNew = fold_partials(rzrr, np), rr = fold_partials(rzrr + np, np);
    double beta = 0;
    if (oldSlot >= 0) {
        double o = sc[oldSlot];
        beta = o != 0 ? rzNew / o : 0.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) sc[newSlot] = rzNew, sc[S_RR] = rr;
    for (long long i = (long long)blockIdx.x * B + threadIdx.x; i < n; i += (long long)gridDim.x * B) p[i] = z[i] + (oldSlot >= 0 ? beta * p[i] : 0.0);
}
// ---- Conformal basis, two-cycle preconditioner: z = kappa C M C r on each half, C = one cycle of the scalar multigrid
// hierarchy on M + eps K (six-channel layout [V][6]: the two halves ride in channels 0 and 1, the others stay zero).
__global__ void k_conformal_pack(const double* __restrict__ r, int V, double* __restrict__ r6) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double* d = r6 + 6 * (size_t)v;
    d[0] = r[v], d[1] = r[v + V], d[2] = d[3] = d[4] = d[5] = 0;
}
__global__ void k_conformal_weight(const double* __restrict__ z6, const double* __restrict__ m0, int V, double* __restrict__ r6) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double* d = r6 + 6 * (size_t)v;
    d[0] = m0[v] * z6[6 * (size_t)v], d[1] = m0[v] * z6[6 * (size_t)v + 1], d[2] = d[3] = d[4] = d[5] = 0;
}
__global__ void k_conformal_unpack(const double* __restrict__ z6, double kappa, int V, double* __restrict__ z) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    z[v] = kappa * z6[6 * (size_t)v], z[v + V] = kappa * z6[6 * (size_t)v + 1];
}
// z -= mean(z) on each half: the projector onto the complement of the operator's null space (the constants of either potential)
__global__ void k_conformal_remove_means(const double* __restrict__ sums, int V, double* __restrict__ z) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    z[v] -= sums[0] / V, z[v + V] -= sums[1] / V;
}
// the diagonal of the stiffness matrix / the trace of a vertex's diagonal block of P^T D P (to balance eps)
__global__ void k_stiffness_diagonal(const int* __restrict__ rowptr, const int* __restrict__ he, const double* __restrict__ stiff, int V, double* __restrict__ out) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double d = 0;
    for (int k = rowptr[v]; k < rowptr[v + 1]; k++)
        if (he[k] < 0) d = stiff[k];
    out[v] = d;
}
__global__ void k_block_trace(const double* __restrict__ blk, int V, double* __restrict__ out) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) out[v] = blk[4 * (size_t)v] + blk[4 * (size_t)v + 2];
}

__global__ void k_residual(const double* __restrict__ b, const double* __restrict__ ax, long long n, double* __restrict__ r) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = b[i] - ax[i];
}

// ------------------------------------------------------------------------- step and triangle field

// z = (P x)_t, out[t] = z^T D_t z (for x . Dt x, VectorField.h:91-93); Connection: z = x_t.
__global__ void k_connection_step_terms(const double* __restrict__ D, const double* __restrict__ x, int T, double* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double z[2] = {x[2 * t], x[2 * t + 1]};
    out[t] = quad(D + 3 * (size_t)t, z, z);
}
__global__ void k_conformal_step_terms(const int* __restrict__ tri, const double* __restrict__ g, const double* __restrict__ D, const double* __restrict__ x, int V,
                                       int T, double* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    ConformalP p;
    conformal_p(g, t, p);
    double z[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int v = tri[3 * t + k];
        double a = x[v], b = x[v + V];
        z[0] += p.pg[k][0] * a + p.pr[k][0] * b, z[1] += p.pg[k][1] * a + p.pr[k][1] * b;
    }
    out[t] = quad(D + 3 * (size_t)t, z, z);
}
// step = (x.b) / (x.Dt x); coeffs += step * x (VectorField.h:93-99).
__global__ void k_vf_update_coeffs(const double* __restrict__ x, const double* __restrict__ scalars, long long n, double* __restrict__ coeffs) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double denom = scalars[SC_STEP_DEN] * scalars[SC_DATA_SCALE], num = scalars[SC_STEP_NUM];
    double step = denom ? num / denom : 0.0;
    if (step) coeffs[i] += x[i] * step;
}
// GetTriangleVectorField, VectorField.h:107-112: tField = P coeffs.
__global__ void k_conformal_field(const int* __restrict__ tri, const double* __restrict__ g, const double* __restrict__ coeffs, int V, int T,
                                  double* __restrict__ tfield) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    ConformalP p;
    conformal_p(g, t, p);
    double z[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int v = tri[3 * t + k];
        double a = coeffs[v], b = coeffs[v + V];
        z[0] += p.pg[k][0] * a + p.pr[k][0] * b, z[1] += p.pg[k][1] * a + p.pr[k][1] * b;
    }
    tfield[2 * t] = z[0], tfield[2 * t + 1] = z[1];
}
__global__ void k_invert(const double* __restrict__ in, int n, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 1.0 / in[i];
}

}  // namespace vfk
}  // namespace mof
