// Spectrum — drop-in command line for the batch part of the reference's sibling tool of the same name (Spectrum/Spectrum.cpp):
// reads a triangle mesh, computes the lowest eigenvectors of the vector Laplacian of the chosen basis (ComputeSpectrum,
// include/Src/VectorLaplacianSpectrum.inl:5-39) and writes them, prolonged to one 2-vector per triangle, to
// eigenvector-001.bin ... in the working directory (Spectrum.cpp:185-189; WriteVector, Src/VectorIO.h:23-31: int count, then
// count x 2 doubles). The reference then opens an OpenGL window to browse them; this build has no viewer and stops there.
//
//     Spectrum --mesh mesh.ply [--vfMode 0|1|2] [--cMode 0|1|2] [--eigenVectors 20]
//
// The eigenproblem runs on the GPU through mof_spectrum (include/mof_b200.h; csrc/spectrum.cu). Flags as in Spectrum.cpp:58-62.
// --eigenVectors is listed by the reference's usage text but missing from its parameter table (:62), so the reference always
// computes 20; here it is honoured. --eLength is parsed and, like in the reference, never used. --edgeMetric (a metric from
// per-face edge lengths, Src/MetricFace.h) is refused: the C ABI takes embedded meshes.
#include <strings.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mof_b200.h"
#include "ply_io.h"

namespace {

void show_usage(const char* ex) {
    printf("Usage %s:\n", ex);
    printf("I/O Parameters: \n");
    printf("\t[--%s <input geometry (.ply)>]\n", "mesh");
    printf("Processing Parameters: \n");
    printf("\t[--%s <number of eigenvector> = %02d]\n", "eigenVectors", 20);
    printf("\t[--%s <subdivide edges up to this diagonal fraction> = %0.3f]\n", "eLength", 0.f);
    printf("\t[--%s <metric from edge length>]\n", "edgeMetric");
    printf("Vector Field Parameters: \n");
    printf("\t[--%s <vector field mode >=%d]\n", "vfMode", 0);
    printf("\t \t [%d] Whitney \n", 0);
    printf("\t \t [%d] Conformal \n", 1);
    printf("\t \t [%d] Connection \n", 2);
    printf("\t[--%s <connection mode >=%d]\n", "cMode", 0);
    printf("\t \t [%d] Projected baricentric \n", 0);
    printf("\t \t [%d] Baricentric dual \n", 1);
    printf("\t \t [%d] Inverse cotangents \n", 2);
}

}  // namespace

int main(int argc, char** argv) {
    std::string mesh;
    int vfMode = 0, cMode = 0, count = 20, device = 0, maxIterations = 20000;
    double tol = 1e-8;
    bool edgeMetric = false;
    for (int i = 1; i < argc; i++) {  // cmdLineParse, CmdLineParser.inl:239-262: `--name value`, names case-insensitive
        const char* a = argv[i];
        if (a[0] != '-' || a[1] != '-') {
            fprintf(stderr, "[WARNING] Parameter name should be of the form --<name>: %s\n", a);
            continue;
        }
        const char* name = a + 2;
        auto value = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (!strcasecmp(name, "mesh")) mesh = value();
        else if (!strcasecmp(name, "vfMode")) vfMode = atoi(value());
        else if (!strcasecmp(name, "cMode")) cMode = atoi(value());
        else if (!strcasecmp(name, "eigenVectors")) count = atoi(value());
        else if (!strcasecmp(name, "eLength")) (void)value();
        else if (!strcasecmp(name, "edgeMetric")) edgeMetric = true;
        else if (!strcasecmp(name, "device")) device = atoi(value());
        else if (!strcasecmp(name, "tolerance")) tol = atof(value());
        else if (!strcasecmp(name, "maxIterations")) maxIterations = atoi(value());
        else {
            fprintf(stderr, "[WARNING] Invalid option: %s\n", a);
            fprintf(stderr, "\t--mesh\n\t--vfMode\n\t--cMode\n\t--eLength\n\t--edgeMetric\n");
        }
    }
    if (mesh.empty()) {
        show_usage(argv[0]);
        return EXIT_FAILURE;
    }
    if (edgeMetric) {
        fprintf(stderr, "[ERROR] --edgeMetric (a metric from per-face edge lengths) is not supported by this build\n");
        return EXIT_FAILURE;
    }
    if (vfMode < 0 || vfMode > 2) {
        printf("ERROR: Unsupported vector field! \n");  // Spectrum.cpp:180
        return 0;
    }
    mof::PlyMesh ply;
    std::string err;
    if (!mof::ply_read(mesh.c_str(), ply, err)) {
        printf("Unable to read %s. Check the file name and format!", mesh.c_str());  // Spectrum.cpp:160
        return 0;
    }
    const size_t V = ply.vertexCount(), T = ply.faceCount();
    std::vector<double> xyz(3 * V);
    for (size_t i = 0; i < 3 * V; i++) xyz[i] = (double)ply.xyz[i];
    std::vector<int> tri(3 * T);
    for (size_t f = 0, at = 0; f < T; at += ply.faceSize[f], f++) {
        if (ply.faceSize[f] != 3) {
            fprintf(stderr, "[ERROR] Polygon is not a triangle: %d != %d\n", ply.faceSize[f], 3);  // PlyReadTriangles, Ply.inl
            return 0;
        }
        for (int k = 0; k < 3; k++) tri[3 * f + k] = ply.faceIndex[at + k];
    }
    mof_ctx* ctx = nullptr;
    if (mof_create(device, nullptr, &ctx) != MOF_OK) {
        fprintf(stderr, "[ERROR] no usable CUDA device: this build has no CPU path\n");
        return EXIT_FAILURE;
    }
    auto ok = [&](int rc) {
        if (rc == MOF_OK) return true;
        fprintf(stderr, "%s\n", mof_last_error(ctx));
        return false;
    };
    mof_params params;
    mof_default_params(&params);
    params.vfMode = vfMode, params.cMode = cMode;
    if (!ok(mof_set_params(ctx, &params)) || !ok(mof_set_mesh(ctx, xyz.data(), (int)V, tri.data(), (int)T))) return 0;
    std::vector<double> values((size_t)count), fields((size_t)count * 2 * T);
    int iterations = 0;
    double residual = 0;
    printf("Solving Eigenvalue Problem\n");
    if (!ok(mof_spectrum(ctx, count, tol, maxIterations, values.data(), fields.data(), &iterations, &residual))) {
        printf("Unable to Compute Laplacian Spectrum \n");  // VectorLaplacianSpectrum.inl:27
        return 0;
    }
    // the summary ComputeEigenvectors prints (EigenvalueSolver.h:117-140), with this solver's counters in ARPACK's places
    printf("Real symmetric eigenvalue problem: A*x - B*x*lambda\n");
    printf("Dimension of the system            : %lld\n", mof_num_coeffs(ctx));
    printf("Number of 'requested' eigenvalues  : %d\n", count);
    printf("Number of 'converged' eigenvalues  : %d\n", count);
    printf("Number of iterations taken         : %d\n\n", iterations);
    printf("Eigenvalues:\n");
    for (int j = 0; j < count; j++) printf("%0.8f \n", values[j]);
    for (int j = 0; j < count; j++) {
        char name[256];
        snprintf(name, sizeof(name), "eigenvector-%03d.bin", j + 1);
        FILE* f = fopen(name, "wb");
        if (!f) {
            fprintf(stderr, "[ERROR] cannot write %s\n", name);
            return EXIT_FAILURE;
        }
        const int n = (int)T;
        fwrite(&n, sizeof(int), 1, f);
        fwrite(fields.data() + (size_t)j * 2 * T, sizeof(double) * 2, T, f);
        fclose(f);
    }
    mof_destroy(ctx);
    fprintf(stderr, "[WARNING] this build has no viewer: %d eigenvectors written to eigenvector-001.bin ...\n", count);
    return EXIT_SUCCESS;
}
