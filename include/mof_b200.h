/* mof_b200.h — C ABI of the B200-native halfway-alignment solver (libmof_b200.so).
 *
 * The reference (fabianprada/MeshOpticalFlow) has no plugin/FFI interface: its solver loop is
 * reached only through process-global statics inside OpticalFlow/OpticalFlow.cpp. This header is the
 * boundary a maintainer would bind instead; every entry point names the reference code it replaces
 * (paths relative to the reference root). INTEGRATION.md shows the call-for-call replacement inside
 * OpticalFlow.cpp.
 *
 * Rules: plain pointers and sizes only; every function returns 0 on success or a negative
 * MOF_E_* code and never throws, exits or prints (mof_last_error gives the reference-style
 * message); the caller owns every host buffer; the context owns all device memory; one context
 * per GPU/stream, no hidden globals. Arithmetic is fp64 with int32 indices, like the reference
 * (_main<double,3>, OpticalFlow.cpp:1115). There is no CPU fallback: without a CUDA device
 * mof_create fails.
 */
#ifndef MOF_B200_H
#define MOF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MOF_OK 0
#define MOF_E_INVALID -1      /* bad argument / call order */
#define MOF_E_CUDA -2         /* CUDA runtime error (message has the cudaError string) */
#define MOF_E_MESH -3         /* non-manifold or open mesh: "[ERROR] Edge is occupied" (FEM.inl:599), "[ERROR] Boundary edge" (FEM.inl:554) */
#define MOF_E_NOCONVERGE -4   /* PCG hit its iteration cap */
#define MOF_E_UNSUPPORTED -5  /* a combination outside the accelerated path (e.g. a partitioned mesh with vfMode 1|2) */

typedef struct mof_ctx mof_ctx;

/* Solver parameters: the reference's command-line flags (OpticalFlow.cpp:56-63) with the defaults of
 * _main (:1062-1069), plus the PCG controls that replace the direct factorisation. */
/* A PCG solve that stagnates above its tolerance but below this TRUE relative residual is accepted (and counted in
 * mof_stats.solvesAboveTolerance); above it the call fails with MOF_E_NOCONVERGE. north_star's flow tolerance is 1e-8. */
#define MOF_ACCEPT_RELRES 1e-4

typedef struct mof_params {
    int iterations;        /* --iterations 10 */
    double sSmooth;        /* --sSmooth 3e-3 (float literal in the reference) */
    double sMultiply;      /* --sMultiply 0.25 */
    double vfSmooth;       /* --vfSmooth: 3e-6 (Whitney), 5e-7 (Conformal), 1e4 (Connection), OpticalFlow.cpp:1067-1069; <= 0 selects the mode's default */
    double vMultiply;      /* --vMultiply 1 */
    double vfSThreshold;   /* --vfSThreshold 1e-8 */
    double dogWeight;      /* --dogWeight 1 (0 disables the DoG normalisation; 0<w<1 blends raw and DoG signals as 6 channels, OpticalFlow.cpp:849-855, 1114) */
    double dogSmooth;      /* --dogSmooth 1e-4 */
    double flowTol;        /* relative residual ||r||/||b|| of the flow-system PCG, default 1e-8 */
    double smoothTol;      /* relative residual of the scalar smoothing PCG, default 1e-10 */
    int maxCgIterations;   /* PCG cap, default 100000 */
    int vfMode;            /* --vfMode 0 Whitney | 1 Conformal | 2 Connection (VectorField.h:3-7); read by mof_set_signals */
    int cMode;             /* --cMode 0 projected barycentric | 1 barycentric dual | 2 inverse cotangent (Connection.inl:1-5) */
    int logSpace;          /* --log: compare log(max(1, x)) * 255 / log(255) (OpticalFlow.cpp:821); the colours advected at the end stay as given; read by mof_set_signals */
} mof_params;

/* Counters since mof_create / mof_reset_stats. */
typedef struct mof_stats {
    long long kernelLaunches;     /* kernels launched by this library */
    long long flowCgIterations;   /* PCG iterations spent in flow solves (VectorField.h:85) */
    long long smoothCgIterations; /* PCG iterations spent in smoothing solves (OpticalFlow.cpp:364, :840) */
    int flowSolves, smoothSolves;
    double lastFlowResidual;      /* relative residual reached by the last flow solve */
    double lastSmoothResidual;
    float flowSolveMs;            /* device time (CUDA events on the context's stream) in flow PCG kernels */
    float smoothSolveMs;
    float advectMs;               /* ... in the advection kernels */
    float setupMs;                /* ... in mof_set_mesh */
    double flowSpmvBytes;         /* algorithmic bytes of ONE flow-system SpMV: 12*nnz + 4*(n+1) + 16*n */
    long long flowRows, flowNnz;
    long long haloEntries;        /* partitioned mesh: vector entries this rank receives per halo exchange (0 otherwise) */
    int solvesAboveTolerance;     /* solves whose TRUE relative residual ended above the requested tolerance: accepted when below
                                     MOF_ACCEPT_RELRES (a stagnated recurrence, after its restarts), an error (MOF_E_NOCONVERGE) otherwise.
                                     0 in every configuration the tests and the bench run; the reference's direct solves reach ~1e-14 */
} mof_stats;

void mof_default_params(mof_params* p);

/* One context per GPU; `stream` is a cudaStream_t (NULL = a stream owned by the context). */
int mof_create(int device, void* stream, mof_ctx** out);
void mof_destroy(mof_ctx* ctx);
const char* mof_last_error(const mof_ctx* ctx);
int mof_set_params(mof_ctx* ctx, const mof_params* p);
int mof_get_stats(mof_ctx* ctx, mof_stats* out);
void mof_reset_stats(mof_ctx* ctx);
int mof_synchronize(mof_ctx* ctx);

/* Mesh setup. Replaces, in WhitneyFlowViewer::Init (OpticalFlow.cpp:787-815, 863-870):
 *   setMetricFromEmbedding / makeUnitArea / setInverseMetric  FEM.inl:1305-1323, 1283-1291, 1363-1369
 *   getEdgeXForms                                              FEM.inl:543-614
 *   scalarMassMatrix(false) / scalarStiffnessMatrix            FEM.inl:1507-1549, 439-496
 *   triangleArea                                               OpticalFlow.cpp:813-814
 *   WhitneyVectorField::Init                                   Whitney.inl:28-180
 * xyz: V x 3 positions; tri: T x 3 vertex indices. Host pointers (the _device variant takes device
 * pointers valid on the context's stream). Resets signals and flow. */
int mof_set_mesh(mof_ctx* ctx, const double* xyz, int V, const int* tri, int T);
int mof_set_mesh_device(mof_ctx* ctx, const double* d_xyz, int V, const int* d_tri, int T);

/* Numbering. The reference's solver applies a fill-reducing ordering of its own inside the factorisation
 * (Eigen::SimplicialLDLT, LinearSolvers.h:277, AMD by default), so it is indifferent to how the PLY file numbers
 * vertices and triangles; a gather-based GPU path is not (1M vertices: 46 ms per UpdateFlow numbered along a space-filling
 * curve, 71 ms in creation order, 83 ms shuffled). mof_set_mesh therefore renumbers a mesh of >= 65 536 vertices whose
 * numbering is not local (mean index span of a triangle, or mean jump between consecutive triangles, above V/16) along the
 * Morton curve of positions / centroids, and mof_set_signals, mof_get_flow and mof_advect_vertices translate at the
 * boundary: the caller only ever sees its own numbering. mode: -1 (default) decide by that measure, 0 never, 1 always;
 * takes effect at the next mof_set_mesh. What stays in the library's numbering: the debug taps (mof_get_csr,
 * mof_get_array, mof_get_coeffs) - mof_get_permutation returns the two orders (new index -> caller's index; identity when
 * nothing was renumbered). The texture-map entries (mof_set_texture_map, mof_build_texture_map) follow the caller's
 * triangle ORDER (first-writer rule of RasterizeTriangle, MeshFlow.inl:281-337) and refuse a renumbered mesh: call
 * mof_set_reorder(ctx, 0) first, as the command line does for --mesh. */
int mof_set_reorder(mof_ctx* ctx, int mode);
int mof_get_permutation(mof_ctx* ctx, int* reordered, int* vertexOrder, int* triangleOrder);

/* The two signals to align, V x channels each (channels must be 3), as in flowData.signals
 * (OpticalFlow.cpp:745-751, 772-779), followed by the difference-of-Gaussians normalisation of
 * OpticalFlow.cpp:822-857 when params.dogWeight > 0. The raw values are kept for
 * mof_advect_vertices. Resets the flow to zero. */
int mof_set_signals(mof_ctx* ctx, const double* a, const double* b, int channels);
int mof_set_signals_device(mof_ctx* ctx, const double* d_a, const double* d_b, int channels);

/* n iterations of UpdateFlow with the weight schedule of IterativeOptimization
 * (OpticalFlow.cpp:424-474, 1037-1043; VectorField::UpdateOpticalFlow, VectorField.h:46-104). The
 * schedule continues across calls; mof_set_signals restarts it. */
int mof_iterate(mof_ctx* ctx, int n);

/* tFlowField (T x 2, OpticalFlow.cpp:280) and the basis coefficients (VectorField.h:20): mof_num_coeffs of them —
 * E for Whitney (one per edge), 2V for Conformal ([potential; co-potential], Conformal.inl:14), 2T for Connection
 * ([2t+r], Connection.inl:25). */
int mof_get_flow(mof_ctx* ctx, double* tField);
int mof_get_coeffs(mof_ctx* ctx, double* coeffs);
int mof_num_edges(mof_ctx* ctx);
long long mof_num_coeffs(mof_ctx* ctx);

/* The reference's sibling tool `Spectrum` (Spectrum/Spectrum.cpp:184-190): the `count` lowest eigenpairs of the vector Laplacian of
 * the basis in params.vfMode / cMode, S x = lambda M x with M = R (g area) P — ComputeSpectrum, include/Src/VectorLaplacianSpectrum.inl:5-39,
 * which factorises S - 1e-8 M and runs ARPACK's shift-invert Lanczos (EigenvalueSolver.h:177-219). Here: LOBPCG on the operators the
 * context already holds (csrc/spectrum.cu), no factorisation. Needs the mesh only; afterwards mof_set_signals has to be called again
 * before an alignment. eigenvalues[count] ascending; fields[count][T][2] = the prolonged eigenvectors P x (what the tool writes to
 * eigenvector-%03d.bin), x normalised to x^T M x = 1 like ARPACK's, sign free; inside a multiple eigenvalue only the span is defined.
 * Converged when every pair has ||S x - lambda M x|| <= tol (||S x|| + max(|lambda|, lambda_count / 1000) ||M x||) - the floor is for the
 * harmonic fields of a surface of genus > 0, whose lambda is 0; MOF_E_NOCONVERGE after maxIterations.
 * count <= 28 (a block of min(32, count + max(4, count/2 + 2)) vectors: the guard keeps the block from ending inside the cluster of the
 * last wanted eigenvalue). */
int mof_spectrum(mof_ctx* ctx, int count, double tol, int maxIterations, double* eigenvalues, double* fields, int* iterations, double* residual);

/* InputGeometryData::flow (OpticalFlow.cpp:482-489): the raw signals resampled along -alpha and
 * 1-alpha of the flow (ResampleSignal, :198-216). outA/outB: V x 3. */
int mof_advect_vertices(mof_ctx* ctx, double alpha, double* outA, double* outB);
int mof_advect_vertices_device(mof_ctx* ctx, double alpha, double* d_outA, double* d_outB);

/* Texel variant. mof_set_texture_map uploads what GetTextureSource produced (MeshFlow.inl:411-467):
 * srcT[W*H] triangle per texel (-1 = uncovered), srcP[W*H][2] barycentric point, triUV[T][6]
 * per-corner texture coordinates, and the two RGB8 textures (top row first, as PNGReadColor
 * returns them). mof_advect_texels is InputTextureData::flow (OpticalFlow.cpp:501-515) with
 * Sample (MeshFlow.inl:66-84): out[s][W*H][3] in the reference's bottom-up texel order; uncovered
 * texels are set to the vertically flipped input (OpticalFlow.cpp:889). */
int mof_set_texture_map(mof_ctx* ctx, int W, int H, const int* srcT, const double* srcP, const double* triUV,
                        const unsigned char* texA, const unsigned char* texB);
int mof_advect_texels(mof_ctx* ctx, double alpha, int bilinear, double* outA, double* outB);
/* The frame sequence of the same (InputTextureData::flow(flowData, frames, ...), OpticalFlow.cpp:517-539 — what the reference's
 * viewer renders as an animation): the texels' sample points are carried along the flow in frames - 1 equal steps of
 * 1 / (frames - 1) — backwards for the first signal, forwards for the second, minimum step 1e-2 * frames — and the texture is
 * fetched after every step. out[s]: [frames][W*H][3]; frame 0, and the uncovered texels of every frame, are the flipped input. */
int mof_advect_texels_frames(mof_ctx* ctx, int frames, int bilinear, double* outA, double* outB);

/* The texture configuration's one-time preparation on the device (the reference runs it serially on the CPU; the
 * outputs are the same, integer for integer and bit for bit: see csrc/texprep_kernels.cu).
 *
 * mof_subdivide is Subdivide (MeshFlow.inl:223-232 over _Subdivide :158-220): every side longer than edgeLength is
 * split at its midpoint, sweep after sweep, until none is left; midpoint vertices are numbered in the order the
 * reference's scan meets their edges. xyz[V][3] single precision like the reference's PlyVertex<float>, tri[T][3],
 * triUV[T][6]. Needs no mesh in place; the result stays in the context until mof_get_subdivision copies it out
 * (xyz[outV][3], tri[outT][3], triUV[outT][6]). edgeLength <= 0 leaves the mesh as it is (OpticalFlow.cpp:714).
 *
 * mof_build_texture_map is GetTextureSource (MeshFlow.inl:411-467) for the mesh in place: RasterizeTriangle (:281-337,
 * first triangle wins except under the rule of :334), padRadius rings of padding (:426-455), RemapSamplePoint through
 * RiemannianMesh::exp (:340-350, FEM.inl:835-899) over the edge transforms the library built. It installs the map and
 * the two textures exactly as mof_set_texture_map would. A ray that misses its triangle makes the reference exit
 * (FEM.inl:889): here MOF_E_MESH with its message, *misses = the number of such texels. mof_get_texture_map reads the
 * installed map back (srcT[W*H], srcP[W*H][2]).
 *
 * mof_sample_textures_to_vertices is SampleTextureToVertices (MeshFlow.inl:252-266, Sample :66-84) of the two installed
 * textures: outA/outB[V][3], each vertex the mean of its wedges' samples (summed in triangle order). */
int mof_subdivide(mof_ctx* ctx, const float* xyz, int V, const int* tri, const double* triUV, int T, double edgeLength, int* outV, int* outT);
int mof_get_subdivision(mof_ctx* ctx, float* xyz, int* tri, double* triUV);
int mof_build_texture_map(mof_ctx* ctx, int W, int H, int padRadius, const double* triUV, const unsigned char* texA, const unsigned char* texB,
                          int* misses);
int mof_get_texture_map(mof_ctx* ctx, int* srcT, double* srcP);
int mof_sample_textures_to_vertices(mof_ctx* ctx, int bilinear, double* outA, double* outB);

/* Debug taps for the parity tests. */
enum {
    MOF_CSR_SCALAR_MASS = 0,      /* flowData.sMass       V x V  (FEM.inl:1548) */
    MOF_CSR_SCALAR_STIFFNESS = 1, /* flowData.sStiffness  V x V  (FEM.inl:1549) */
    MOF_CSR_WHITNEY_SMOOTH = 2,   /* vf->smoothOperator   E x E  (Whitney.inl:179) */
    MOF_CSR_FLOW_SYSTEM = 3       /* opticalFlowMatrix of the last iteration (VectorField.h:67) */
};
/* rows/nnz query, then copy out with columns ascending in every row. */
int mof_csr_size(mof_ctx* ctx, int which, int* rows, long long* nnz);
int mof_get_csr(mof_ctx* ctx, int which, int* rowptr, int* col, double* val);

enum {
    MOF_ARR_METRIC = 0,          /* T x 3 (g00,g01,g11) after makeUnitArea */
    MOF_ARR_AREA = 1,            /* T */
    MOF_ARR_OPPOSITE = 2,        /* 3T int32   EdgeXForm::oppositeEdge */
    MOF_ARR_XFORM_LINEAR = 3,    /* 3T x 4 row-major */
    MOF_ARR_XFORM_CONSTANT = 4,  /* 3T x 2 */
    MOF_ARR_REDUCED_EDGE = 5,    /* 3T int32   reducedEdgeIndex (Whitney.inl:33) */
    MOF_ARR_EXPANDED_EDGE = 6,   /* E int32    expandedEdgeIndex */
    MOF_ARR_POSITIVE_EDGE = 7,   /* 3T int32   positiveOrientedEdge as 0/1 */
    MOF_ARR_PROLONGATION = 8,    /* T x 3 x 2  P entries: row 2t+r, k-th edge of t at [t][k][r] (Whitney.inl:80-83) */
    MOF_ARR_SIGNALS = 9,         /* V x 6      flowData.signals after DoG: (A rgb, B rgb) per vertex */
    MOF_ARR_SMOOTHED = 10,       /* V x 6      last iteration's smoothed signals (OpticalFlow.cpp:435) */
    MOF_ARR_RESAMPLED = 11,      /* V x 6      last iteration's resampled signals (:439) */
    MOF_ARR_DATA_TERM = 12,      /* T x 3      (d00,d01,d11) (:395-421) */
    MOF_ARR_DATA_RHS = 13,       /* T x 2 */
    MOF_ARR_FLOW_RHS = 14,       /* mof_num_coeffs   scaled R*rhs (VectorField.h:53,60) */
    MOF_ARR_FLOW_SOLUTION = 15,  /* mof_num_coeffs   solution of the last flow system (VectorField.h:85) */
    MOF_ARR_SIGNALS_RAW = 16,    /* V x 6      6-channel blend only: the (1-w)*raw half of flowData.signals; MOF_ARR_SIGNALS is the w*DoG half */
    MOF_ARR_RESAMPLED_RAW = 17   /* V x 6      6-channel blend only: that half of the last iteration's resampled signals */
};
/* bytes needed for one array (0 if not available yet), then copy out. */
long long mof_array_bytes(mof_ctx* ctx, int which);
int mof_get_array(mof_ctx* ctx, int which, void* out);

/* Stand-alone solver entry for tests and profiling: Jacobi-PCG on a caller-supplied SPD CSR system
 * (host pointers), NRHS = 1. Returns iterations in *iters and the relative residual in *relres. */
int mof_pcg_solve_csr(mof_ctx* ctx, int n, const int* rowptr, const int* col, const double* val, const double* b, double* x,
                      double tol, int maxIters, int* iters, double* relres);
/* Times `reps` launches of the flow-system SpMV kernel (y = A d fused with d.y) on the context's
 * current flow matrix; returns average ms per launch. Used by bench.py for the roofline line. */
int mof_time_flow_spmv(mof_ctx* ctx, int reps, float* msPerLaunch);
/* The same for the other kernels of a PCG iteration and the walk, each launched `reps` times on its own with the solver's own
 * grid and the context's current operators (CUDA events on the context's stream): average microseconds per launch and the
 * ALGORITHMIC bytes one launch moves (every array touched once; 0 where trip counts are data dependent). bench.py builds its
 * per-kernel roofline table from this. Needs a completed mof_iterate with the Whitney basis; scratch vectors of the solvers
 * are overwritten (results already computed are not). */
enum mof_kernel_id {
    MOF_K_FLOW_SPMV = 0,          /* k_spmv_dot: fp64 y = A p fused with p.y (the PCG's matrix product) */
    MOF_K_FLOW_FINE_SWEEP = 1,    /* k_fine_apply_flow<float>: one damped-Jacobi sweep of the cycle on the fp32 copy, fused with r.z */
    MOF_K_FLOW_UPDATE = 2,        /* k_update_xr: x += alpha p, r -= alpha q, r.r, first sweep of the next cycle */
    MOF_K_FLOW_RESTRICT = 3,      /* k_restrict_flow: fine residual -> level-1 cells (+ their first sweep) */
    MOF_K_FLOW_PROLONG = 4,       /* k_prolong_flow */
    MOF_K_FLOW_DIRECTION = 5,     /* k_direction: p = z + beta p */
    MOF_K_FLOW_LEVEL1 = 6,        /* k_coarse_apply<9,3,.> on the largest coarse level: 27-point stencil of 3x3 blocks */
    MOF_K_SCALAR_SPMV = 7,        /* fp64 six-channel product of the smoothing PCG, fused with p.q */
    MOF_K_SCALAR_FINE_SWEEP = 8,  /* fp32 six-channel sweep of the smoothing cycle */
    MOF_K_SCALAR_UPDATE = 9,
    MOF_K_SCALAR_LEVEL1 = 10,
    MOF_K_WALK = 11,              /* k_walk_sample along the current flow (bytes = 0: data-dependent trip counts) */
    MOF_K_COUNT = 12
};
int mof_time_kernel(mof_ctx* ctx, int which, int reps, double* usPerLaunch, double* algorithmicBytes);

/* ---- one mesh over several GPUs (BASELINE.json configs[4]: "vertex-partitioned, NCCL-over-NVLink halo exchange for
 * SpMV and allreduce for the CG dot products"). The reference has no counterpart: its solve is one Eigen
 * factorisation (LinearSolvers.h:360-391). One process per GPU; rank 0 obtains an id, the host side distributes it
 * (bench.py / the tests use torch.distributed for that), every rank calls mof_dist_init BEFORE mof_set_mesh, and from
 * then on every rank makes the SAME calls with the SAME inputs: the flow solves are then row-partitioned across the
 * ranks (halo exchange + all-reduce on the context's stream), everything else is replicated, and every rank ends
 * with the full result. world == 1 is allowed (same code path, no communication). */
int mof_dist_unique_id(unsigned char id128[128]);
int mof_dist_init(mof_ctx* ctx, int world, int rank, const unsigned char id128[128]);

#ifdef __cplusplus
}
#endif
#endif
